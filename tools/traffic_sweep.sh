#!/bin/bash
# DRAM traffic of ONE main-pass launch of the fused kernel at the bench size for several
# (DL_FUSED_G, DL_FUSED_TT) settings: usage traffic_sweep.sh "G=148 TT=64" "G=74 TT=128" ...
export PYTHONPATH=.
export DL_FUSED_NORESCUE=1
FULL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
$FULL > gpurun_out/ts_plain.log 2>&1 || exit 1
for cfg in "$@"; do
  eval "$cfg"
  echo "=== G=$G TT=$TT"
  DL_FUSED_G=$G DL_FUSED_TT=$TT ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    --clock-control none -k regex:fused_like_kernel -s 3 -c 1 --csv $FULL 2>/dev/null | grep fused_like | awk -F'","' '{print $13, $15}' | tr -d '"'
done
