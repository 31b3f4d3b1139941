"""Turn the ncu CSV of tools/prof_r2.sh (metrics of ONE main-pass launch of the fused kernel at
the bench shape) into profiles/r2_fused_traffic.json, the record bench.py reports as
`roofline.traffic`.  Run HERE right after the GPU call, before touching the kernel sources:
    python tools/traffic_record.py gpurun_out/<tag>_fused_full_traffic.csv [nO nT nz nB]"""
import csv
import datetime
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

path = sys.argv[1]
shape = [int(x) for x in sys.argv[2:6]] if len(sys.argv) >= 6 else [10000, 10000, 1500, 6]
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and "fused_like_kernel" in r[4]]
main = [r for r in rows if r[4].rstrip(")").split(",")[-1].strip().startswith("0>")]
rec = {}
for r in main:
    rec.setdefault(r[12], float(r[14]))
out = {"kernel": main[0][4], "grid": main[0][8], "shape": shape,
       "dram_bytes_read": rec["dram__bytes_read.sum"],
       "dram_bytes_write": rec["dram__bytes_write.sum"],
       "l2_hit_rate_pct": rec.get("lts__t_sector_hit_rate.pct"),
       "duration_ns_under_ncu": rec.get("gpu__time_duration.sum"),
       "fp64_pipe_active_pct": rec.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
       "warp_instructions": rec.get("smsp__inst_executed.sum"),
       "kernel_source_hash": bench.kernel_source_hash(),
       "when": datetime.date.today().isoformat(), "source_csv": os.path.basename(path)}
json.dump(out, open(bench.TRAFFIC_RECORD, "w"), indent=1)
print(json.dumps(out, indent=1))
