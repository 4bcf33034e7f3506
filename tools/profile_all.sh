#!/bin/bash
# Reproduces the measurements under profiles/ on a B200 box (run from the repo root, e.g. through gpurun).
# One GPU unless stated.  Numbers printed by a run under ncu are never bench values.
set -u
mkdir -p gpurun_out
# 1. parity + default bench line (value, e2e, roofline, cpu_baseline)
python -m pytest tests -m gpu -x -q | tail -2
python bench.py > gpurun_out/bench.json
# 2. every launch with its device time and DRAM bytes (cold-cache, serialised: compare shares)
ncu --kernel-name-base demangled -k regex:psacb200 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 78 -c 78 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null
# 3. the dominant kernel, full set, with source
ncu --set full --clock-control none --import-source on -k regex:radix_scatter_seg_kernel -s 5 -c 1 -o gpurun_out/prof_scatter \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null
python tools/ncu_summary.py gpurun_out/launches.csv gpurun_out/prof_scatter.ncu-rep
# 4. micro-benchmarks behind the design decisions
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro_rank tools/micro_rank.cu && ./tools/micro_rank
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/atoms_order tools/atoms_order.cu && ./tools/atoms_order
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DPSAC_PHASE_PROFILE --expt-relaxed-constexpr -o tools/bench_pass tools/bench_pass.cu && ./tools/bench_pass
# 5. sharded construction (N GPUs of one box): parity on 13 cases, then the weak-scaling bench line
#    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/sharded_worker.py
#    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N --steps 3 --warmup 3 [--lcp]
