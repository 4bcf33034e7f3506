#!/bin/bash
# Final round-2 measurements behind profiles/r2b_* (run from the repo root on a B200 box through gpurun; one GPU).
# Of this script the bench commands and the ncu LAUNCH LIST were run for the committed r2b files (each in its own gpurun call);
# the three `--set full` captures below were NOT run any more (the round's GPU budget was spent): the full captures under
# profiles/r2_* are of kernels that did not change afterwards (segmented scatter, pass-1 scatter, ISA scatter).
# Numbers printed by a run under ncu are never bench values.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2b_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err
python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_config1.json 2>> gpurun_out/r2b_bench_1gpu.err
python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_config3.json 2>> gpurun_out/r2b_bench_1gpu.err
python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_config5_1gpu.json 2>> gpurun_out/r2b_bench_1gpu.err
python bench.py --config 6 --steps 5 --warmup 3 > gpurun_out/r2b_bench_gsa.json 2>> gpurun_out/r2b_bench_1gpu.err
# launch list of the default bench command (cold-cache, serialised: compare shares)
ncu --kernel-name-base demangled -k regex:psacb200 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
# full captures: the dominant scatter kernel, the apply phase of the heads kernel (transposed 64-bit LCP stores), the staged tree tile kernel
ncu --set full --clock-control none --import-source on -k regex:radix_scatter_seg_kernel -s 5 -c 1 -o gpurun_out/r2b_prof_scatter_seg \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:heads_kernel -s 3 -c 1 -o gpurun_out/r2b_prof_heads \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:suffix_tree_tile_kernel -s 1 -c 1 -o gpurun_out/r2b_prof_tree_tile \
    python bench.py --config 5 --log2n 28 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2b_launches.csv gpurun_out/r2b_prof_scatter_seg.ncu-rep gpurun_out/r2b_prof_heads.ncu-rep gpurun_out/r2b_prof_tree_tile.ncu-rep > gpurun_out/r2b_ncu_summary.txt 2>&1
ls -la gpurun_out/r2b_* | awk '{print $5, $9}'
