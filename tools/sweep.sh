#!/bin/bash
# usage: sweep.sh "defs1" "defs2" ...   (each a string of -D flags)
for defs in "$@"; do
  echo "=== $defs"
  DL_NVCC_DEFS="$defs" python -m delight_b200.build --force --verbose 2>&1 | grep -E "fused_like_kernelILi6ELb0" -A3 | grep -E "Used|spill" | tr '\n' ' '; echo
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python tools/brief.py
done
python -m delight_b200.build --force > /dev/null
