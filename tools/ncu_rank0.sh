#!/bin/bash
# torchrun child wrapper: rank 0 runs under ncu (per-kernel metrics of OUR kernels only; NCCL kernels run once, un-profiled),
# the other ranks run plain.  usage: torchrun ... --no-python tools/ncu_rank0.sh <csv> <script> [args]
CSV=$1; shift
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_atom.sum \
    --clock-control none --csv --log-file "$CSV" -k regex:"select_words|radix_scatter|tile_hist|heads_kernel|isa_scatter|pull_sa|digit_hist|round_keys|resolve_kernel|convert" python "$@"
else
  exec python "$@"
fi
