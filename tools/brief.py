import sys, json
last_dbg=None
# usage: brief.py FILE   or   ... | brief.py   (never waits on a terminal)
src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for line in src:
    line=line.strip()
    if line.startswith('[dbgwait]'):
        last_dbg=line
    if line.startswith('{'):
        d=json.loads(line)
        r=d.get("roofline") or {}
        print("value=%.0f gal/s  step=%.1f ms  kernel=%.1f ms  frac=%.3f  e2e=%.0f  launches=%s clocks=%s" % (d["value"], d["ms_per_step"], r.get("kernel_ms",0), r.get("frac") or 0, d["e2e"]["value"], d.get("gpu_launches"), (d.get("clocks") or {}).get("sm_mhz")))
if last_dbg: print(last_dbg)
