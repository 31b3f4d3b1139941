#!/bin/bash
# usage: sweep_prebuilt.sh variants/lib_a.so variants/lib_b.so ...
# Runs the bench once per prebuilt library variant (built here with DL_NVCC_DEFS=... python -m
# delight_b200.build --force and copied aside), so no GPU-box time goes into compiling.
cp delight_b200/libdelight_b200.so /tmp/lib_default.so
for lib in "$@"; do
  echo "=== $lib"
  cp "$lib" delight_b200/libdelight_b200.so; touch delight_b200/libdelight_b200.so
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python tools/brief.py
  if [ -n "$SWEEP_TEST" ]; then python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused" 2>&1 | tail -1; fi
done
cp /tmp/lib_default.so delight_b200/libdelight_b200.so; touch delight_b200/libdelight_b200.so
