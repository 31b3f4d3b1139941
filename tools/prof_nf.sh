#!/bin/bash
# ncu --set full of ONE main-pass launch of the fused kernel for a given band count (tools/nf_scan.py
# shape: 2048 targets x 1500 training x 1500 z): usage prof_nf.sh <nf> <tag>
export PYTHONPATH=.
export DL_FUSED_NORESCUE=1
export NF_LIST=$1
CMD="python tools/nf_scan.py"
$CMD > gpurun_out/$2_plain_nf$1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_like_kernel -s 3 -c 1 \
    -f -o gpurun_out/$2_fused_nf$1 $CMD > gpurun_out/$2_ncu_nf$1.log 2>&1
cat gpurun_out/$2_plain_nf$1.log
