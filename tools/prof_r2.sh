#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; outputs in gpurun_out/):
#   1. ncu --set full of ONE main-pass launch of the fused kernel on a reduced case
#   2. DRAM traffic / L2 hit rate / duration of ONE main-pass launch at the bench size
#   3. launch list (device time of every kernel) of a short default bench run
# Every ncu command is preceded by the same command without ncu (must exit 0).
set -x
export PYTHONPATH=.
export DL_FUSED_NORESCUE=1     # two fused launches per call: sampling pass, main pass
SMALL="python bench.py --steps 1 --warmup 3 --targets 2048 --train 2000 --no-cpu-baseline --no-extras"
FULL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
TAG=${1:-r2}
$SMALL > gpurun_out/${TAG}_plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_like_kernel -s 7 -c 1 \
    -f -o gpurun_out/${TAG}_fused_small $SMALL > gpurun_out/${TAG}_ncu_small.log 2>&1
$FULL > gpurun_out/${TAG}_plain_full.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:fused_like_kernel -s 2 -c 2 --csv \
    --log-file gpurun_out/${TAG}_fused_full_traffic.csv $FULL > gpurun_out/${TAG}_ncu_full.log 2>&1
unset DL_FUSED_NORESCUE
LIST="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$LIST > gpurun_out/${TAG}_plain_list.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $LIST > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -2 gpurun_out/${TAG}_plain_full.log
# 4. the other kernels (subsystems 2 and 3, drop-in operator, epilogue): selected counters
SMALLK="python tools/prof_small.py"
$SMALLK > gpurun_out/${TAG}_plain_smallk.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size \
    --clock-control none -k regex:'kernel_grid|kernelparts|gp_fit|gp_predict|topk|loglike|expnorm|compressed|metrics_kernel|ci_levels|finalize|prep_targets|anchor|zprior|relayout' \
    --csv --log-file gpurun_out/${TAG}_small_kernels.csv $SMALLK > gpurun_out/${TAG}_ncu_smallk.log 2>&1
cat gpurun_out/${TAG}_plain_smallk.log
