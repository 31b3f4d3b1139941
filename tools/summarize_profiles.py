"""Turn the raw outputs of tools/prof_r2.sh / tools/prof_nf.sh (gpurun_out/<tag>_*) into the
committed text summaries under profiles/ (run HERE, ncu is installed without a GPU):

    python tools/summarize_profiles.py <tag>

    profiles/r2_fused_like_ncu_full_summary.txt   ncu --set full of one main-pass launch (reduced case)
    profiles/r2_fused_nf21_ncu_summary.txt        the same for 21 bands (configuration 1, two targets jointly)
    profiles/r2_launches_summary.txt              device time per kernel of a default bench run
    profiles/r2_small_kernels.txt                 counters of every other kernel of the library
    profiles/r2_sass_excerpt.txt                  SASS of the hot loop (TMA issue + one cell group)
"""
import collections
import csv
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
PROF = os.path.join(REPO, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2d"

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
       "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "sm__warps_active.avg.per_cycle_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
       "dram__bytes_read.sum", "dram__bytes_write.sum",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                          text=True).stdout


def full_summary(rep, cells, title, dst):
    raw = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    h, units, v = raw[0], raw[1], raw[2]
    lines = [title, "", "Kernel: " + v[h.index("Kernel Name")]]
    for k in RAW:
        if k in h:
            lines.append("%-86s %-10s %s" % (k, units[h.index(k)], v[h.index(k)]))
    src = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv", "--print-source",
                               "sass"]).splitlines()))
    hdr, data = src[1], src[2:]
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = sum(int(r[ia]) for r in data)
    ts = sum(int(r[isamp]) for r in data)
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += int(r[ia])
        samp[op] += int(r[isamp])
    iters = cells / 64.0
    lines += ["", "instructions executed by opcode (warp level, SASS source page; %.0f per "
              "warp-iteration of 2 x 32 cells):" % (tot / iters)]
    for op, c in ops.most_common(22):
        lines.append("  %-10s %5.1f %% of executed  %7.1f per warp-iteration  %5.1f %% of samples"
                     % (op, 100.0 * c / tot, c / iters, 100.0 * samp[op] / ts))
    fp64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
    lines.append("  FP64-pipe instructions: %.1f per warp-iteration = %.1f per cell"
                 % (fp64 / iters, fp64 / iters / 2))
    lines.append("  local memory (spills): LDL %.1f, STL %.1f per warp-iteration"
                 % (ops["LDL"] / iters, ops["STL"] / iters))
    stalls = [c for c in hdr if c.startswith("stall_") and "(" not in c]
    t = collections.Counter()
    for r in data:
        for c in stalls:
            t[c] += int(r[hdr.index(c)] or 0)
    s = sum(t.values())
    lines += ["", "warp-state sampling over the whole kernel (%d samples):" % s]
    for c, n in t.most_common(10):
        lines.append("  %-26s %5.1f %%" % (c, 100.0 * n / s))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)
    return data, hdr


def launches(csvfile, dst):
    rows = [r for r in csv.reader(open(csvfile)) if len(r) > 14 and r[0].isdigit()]
    by = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        if "fused_like_kernel" in r[4]:
            name = r[4].split("(")[0].replace("void ", "")
        d = by.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += float(r[14])
    tot = sum(d[1] for d in by.values())
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none   python bench.py --steps 2 "
             "--warmup 3 --no-cpu-baseline --no-extras   (plain run exited 0 first)",
             "per-launch times under ncu are cold-cache and serialised: compare SHARES", "",
             "%-58s %8s %14s %8s" % ("kernel", "launches", "total ms", "share")]
    for k, (n, ns) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-58s %8d %14.3f %7.2f%%" % (k[:58], n, ns / 1e6, 100 * ns / tot))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


def small_kernels(csvfile, plain, dst):
    rows = [r for r in csv.reader(open(csvfile)) if len(r) > 14 and r[0].isdigit()]
    by = collections.OrderedDict()
    for r in rows:
        by.setdefault((r[0], r[4]), {})[r[12]] = r[14]
    seen = set()
    lines = ["ncu --metrics (time, dram bytes, sm throughput, fp64 pipe, occupancy) of "
             "`python tools/prof_small.py` (every kernel of the library except the fused one; "
             "first launch of each kernel shown)", "",
             "%-44s %10s %9s %9s %8s %6s %6s %5s %9s" % ("kernel", "time us", "read MB", "write MB",
                                                          "DRAM GB/s", "sm %", "fp64 %", "regs",
                                                          "grid")]
    for (i, k), m in by.items():
        name = re.sub(r"\(.*", "", k).replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        t = float(m["gpu__time_duration.sum"]) / 1e3
        rd, wr = float(m["dram__bytes_read.sum"]) / 1e6, float(m["dram__bytes_write.sum"]) / 1e6
        lines.append("%-44s %10.1f %9.1f %9.1f %8.0f %6.1f %6.1f %5s %9s" % (
            name[:44], t, rd, wr, (rd + wr) / t * 1e3 if t else 0,
            float(m["sm__throughput.avg.pct_of_peak_sustained_elapsed"]),
            float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
            m["launch__registers_per_thread"], m["launch__grid_size"]))
    lines += ["", "CUDA-event times of the same script without ncu:"]
    lines += ["  " + ln.rstrip() for ln in open(plain) if " ms" in ln or "bytes" in ln]
    lines += ["", "non-fused likelihood (loglike_kernel + expnorm_kernel): algorithmic bytes / "
              "event time against the measured HBM peak of MEASURED_PEAKS.json (6546 GB/s)."]
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


def sass_excerpt(dst):
    obj = os.path.join(REPO, "delight_b200", "csrc", "like_fused_t0.o")
    fun = "_ZN2dl17fused_like_kernelILi6ELb0ELb0ELi0EEEvNS_11FusedParamsE"
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], stdout=subprocess.PIPE,
                         text=True).stdout
    ins = [re.sub(r"\s*/\*[0-9a-f]{16}\*/\s*$", "", ln).strip() for ln in out.splitlines()
           if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln)]
    ins = [re.sub(r"^/\*([0-9a-f]+)\*/\s*", r"\1  ", i) for i in ins]
    cnt = collections.Counter(re.sub(r"^[0-9a-f]+\s+(@!?U?P\d+\s+)?", "", i).split(" ")[0].split(".")[0]
                              for i in ins)
    tma = [k for k, i in enumerate(ins) if "UBLKCP" in i]
    first_tma = tma[0]
    # start of the LAST run of bulk copies (= the issue inside the step loop; earlier runs are
    # the prologue's), shown from the barrier poll / elect before it
    runs = [k for n, k in enumerate(tma) if n == 0 or k - tma[n - 1] > 40]
    loop_tma = runs[-1] - 95
    first_mufu = next(k for k, i in enumerate(ins) if "MUFU.RCP64H" in i)
    lines = ["cuobjdump -sass -fun fused_like_kernel<6,false,false,MAIN> delight_b200/csrc/like_fused_t0.o",
             "(sm_100a; %d instructions; opcode counts of the whole kernel: %s)" % (
                 len(ins), ", ".join("%s %d" % kv for kv in cnt.most_common(14))), "",
             "---- TMA producer inside the step loop: whole-warp mbarrier poll, elect.sync leader, "
             "expect_tx arrive, 14 bulk copies (6 band means, 6 band variances, prior row, "
             "trapezoid-weight row) issued back to back from uniform registers"]
    lines += ins[loop_tma:loop_tma + 190]
    lines += ["", "---- consumer: mbarrier wait on the full barrier, then the first cell group "
              "(pass 0, MUFU.RCP64H seeds + Newton steps, pass 1 in band pairs, MUFU.RSQ64H, "
              "table exp, max tracking, PDF / evidence accumulation)"]
    lines += ins[first_mufu - 60:first_mufu + 230]
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst, "(first UBLKCP at %d)" % first_tma)


if __name__ == "__main__":
    p = lambda f: os.path.join(OUT, "%s_%s" % (tag, f))
    if os.path.exists(p("fused_small.ncu-rep")):
        full_summary(p("fused_small.ncu-rep"), 2048 * 2000 * 1536,
                     "ncu --set full --clock-control none --import-source on -k regex:fused_like_kernel"
                     " -s 7 -c 1\ncommand: python bench.py --steps 1 --warmup 3 --targets 2048 "
                     "--train 2000 --no-cpu-baseline --no-extras  (DL_FUSED_NORESCUE=1; the same "
                     "command exited 0 without ncu first)\ncells in this launch: 2048 x 2000 x "
                     "1500 (1536 padded) = 6.3e9, main pass of the round-2 kernel",
                     os.path.join(PROF, "r2_fused_like_ncu_full_summary.txt"))
    if os.path.exists(p("fused_nf21.ncu-rep")):
        full_summary(p("fused_nf21.ncu-rep"), 2048 * 1500 * 1536,
                     "ncu --set full --clock-control none --import-source on -k regex:fused_like_kernel"
                     " -s 3 -c 1\ncommand: NF_LIST=21 python tools/nf_scan.py  (2048 targets x "
                     "1500 training x 1500 z x 21 bands; configuration 1: 16 warps, 64-redshift "
                     "stages, two targets jointly)",
                     os.path.join(PROF, "r2_fused_nf21_ncu_summary.txt"))
    if os.path.exists(p("launches.csv")):
        launches(p("launches.csv"), os.path.join(PROF, "r2_launches_summary.txt"))
    if os.path.exists(p("small_kernels.csv")):
        small_kernels(p("small_kernels.csv"), p("plain_smallk.log"),
                      os.path.join(PROF, "r2_small_kernels.txt"))
    sass_excerpt(os.path.join(PROF, "r2_sass_excerpt.txt"))
