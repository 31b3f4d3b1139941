#!/bin/bash
# Parts 1 and 3 of prof_r2.sh only (fused kernel: ncu full set on the reduced case + launch list of a
# short default bench run); for a kernel change that leaves the other kernels alone.
export PYTHONPATH=.
TAG=${1:-r2}
SMALL="python bench.py --steps 1 --warmup 3 --targets 2048 --train 2000 --no-cpu-baseline --no-extras"
export DL_FUSED_NORESCUE=1
timeout 300 $SMALL > gpurun_out/${TAG}_plain_small.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_like_kernel -s 7 -c 1 \
    -f -o gpurun_out/${TAG}_fused_small $SMALL > gpurun_out/${TAG}_ncu_small.log 2>&1
unset DL_FUSED_NORESCUE
LIST="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $LIST > gpurun_out/${TAG}_plain_list.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $LIST > gpurun_out/${TAG}_ncu_list.log 2>&1
ls -la gpurun_out/${TAG}_*
