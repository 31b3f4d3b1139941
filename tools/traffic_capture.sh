#!/bin/bash
# ncu counters of ONE main-pass launch of the fused kernel at the bench size with the default
# settings -> gpurun_out/<tag>_fused_full_traffic.csv (then: python tools/traffic_record.py ...)
export PYTHONPATH=.
export DL_FUSED_NORESCUE=1
TAG=${1:-r2}
FULL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
$FULL > gpurun_out/${TAG}_plain_full.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:fused_like_kernel -s 2 -c 2 --csv \
    --log-file gpurun_out/${TAG}_fused_full_traffic.csv $FULL > gpurun_out/${TAG}_ncu_full.log 2>&1
grep -c fused_like gpurun_out/${TAG}_fused_full_traffic.csv
