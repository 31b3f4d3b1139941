#!/bin/bash
# usage: env_sweep.sh "VAR=val VAR2=val" "..." : runs the bench once per environment setting
# (DL_FUSED_G, DL_FUSED_TT, DL_FUSED_STAGES, DL_FUSED_NSAMPLE are read by dl_like_fused_dev)
for e in "$@"; do
  echo "=== $e"
  env $e python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python tools/brief.py
done
