"""FP64 FMA peak of the device (dl_fp64_peak: dependent DFMA chains on all SMs) with the clocks
seen during the measurement -> one JSON line (kept as profiles/r2_fp64_peak.json)."""
import ctypes
import json
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402
import bench  # noqa: E402
from delight_b200 import _lib  # noqa: E402

L = _lib.lib()
torch.cuda.set_device(0)
v = ctypes.c_double(0)
L.dl_fp64_peak(3, ctypes.byref(v))
s = bench.ClockSampler(0).start()
vals = []
t0 = time.time()
while time.time() - t0 < 4.0:
    L.dl_fp64_peak(20, ctypes.byref(v))
    vals.append(v.value)
clocks = s.stop()
sm = torch.cuda.get_device_properties(0).multi_processor_count
print(json.dumps({"fp64_tflops_best": max(vals), "fp64_tflops_median": sorted(vals)[len(vals) // 2],
                  "runs": len(vals), "clocks": clocks, "sm_count": sm,
                  "datasheet_derived": "%d SM x 64 FMA/clk x 2 flop x 1.965 GHz = %.1f TFLOP/s"
                                       % (sm, sm * 64 * 2 * 1.965e-3),
                  "gpu": torch.cuda.get_device_name(0)}))
