#!/bin/bash
# usage: scan_prebuilt.sh variants/lib_a.so ...   (like sweep_prebuilt.sh, for the other band counts)
# Per prebuilt library: kernel time vs number of bands (tools/nf_scan.py), the G10-style catalogue
# with per-target band lists (tools/het_scan.py) and, with SWEEP_TEST=1, the fused parity tests.
cp delight_b200/libdelight_b200.so /tmp/lib_default.so
for lib in "$@"; do
  echo "=== $lib"
  cp "$lib" delight_b200/libdelight_b200.so; touch delight_b200/libdelight_b200.so
  NF_LIST=${NF_LIST:-4,5,6,7,8,9,13,16,21,28} timeout 300 python tools/nf_scan.py 2>&1 | grep "^nf="
  timeout 300 python tools/het_scan.py 2>&1 | grep "band sets"
  if [ -n "$SWEEP_TEST" ]; then timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused" 2>&1 | tail -1; fi
done
cp /tmp/lib_default.so delight_b200/libdelight_b200.so; touch delight_b200/libdelight_b200.so
