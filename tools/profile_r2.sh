#!/bin/bash
# Round-2 measurements behind profiles/r2_* (run from the repo root on a B200 box, e.g. through gpurun; one GPU).
# Numbers printed by a run under ncu are never bench values.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r2_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config1.json 2>> gpurun_out/r2_bench_1gpu.err
python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config5_1gpu.json 2>> gpurun_out/r2_bench_1gpu.err
# launch list of the default bench command (cold-cache, serialised: compare shares)
ncu --kernel-name-base demangled -k regex:psacb200 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
# full captures: the dominant scatter kernel, digit pass 1 (keys cut from the packed text), the ISA scatter
ncu --set full --clock-control none --import-source on -k regex:radix_scatter_seg_kernel -s 5 -c 1 -o gpurun_out/r2_prof_scatter_seg \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"radix_scatter_kernel.*TextSrc" -s 1 -c 1 -o gpurun_out/r2_prof_pass1 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:isa_scatter_kernel -s 1 -c 1 -o gpurun_out/r2_prof_isa_scatter \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_launches.csv gpurun_out/r2_prof_scatter_seg.ncu-rep gpurun_out/r2_prof_pass1.ncu-rep gpurun_out/r2_prof_isa_scatter.ncu-rep > gpurun_out/r2_ncu_summary.txt 2>&1
# sharded code path on one rank (PSACB200_FORCE_SHARDED): per-kernel metrics of the selection / exchange / heads kernels
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/r2_sharded1_launches.csv -k regex:"select_words|radix_scatter|tile_hist|heads_kernel|isa_scatter|pull_sa|digit_hist" \
    python tools/profile_sharded1.py 28 > /dev/null 2>&1
python tools/ncu_kernels.py gpurun_out/r2_sharded1_launches.csv 0.5 > gpurun_out/r2_sharded1_kernels.txt 2>&1
