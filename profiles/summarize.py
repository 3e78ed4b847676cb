"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries in profiles/.

  python profiles/summarize.py r1        # reads gpurun_out/launches_r1.csv, prof_r1_top.ncu-rep
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w")
P = lambda *a: print(*a, file=out)

# ---- launch list: per-kernel share of one analysis step (cold-cache, serialised)
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv"))))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
agg = collections.OrderedDict()
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
tot = sum(v[1] for v in agg.values())
P(f"# ncu summaries, round {tag}\n")
P("Workload: `python profiles/run_step.py cfg3 2` = bench.py's cfg3 (1080p, 129 frames, GOP 32, block 16,")
P("search 16, quarter-pel), second (warm) analysis step.  Launch list:")
P("`ncu --metrics gpu__time_duration.sum --clock-control none -s 737 -c 737` (times are cold-cache and")
P("serialised: compare shares, not absolutes).\n")
P(f"One step = {sum(v[0] for v in agg.values())} kernel launches, {tot:.1f} ms summed under ncu.\n")
P("| kernel | launches | ms (ncu) | share |")
P("|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    P(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")

# ---- full captures of the top kernels
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_top.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"),
            ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
            ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("launch__registers_per_thread", "registers/thread"),
            ("smsp__inst_executed.sum", "warp instructions"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block")]
    idx = [(h.index(m), n) for m, n in want if m in h]
    units = rr[1]
    P("\n## `ncu --set full --clock-control none` of the level-1 launches of the top kernels\n")
    P("(first temporal level of the warm step: 64 frame pairs; `-k regex:k_ll_residue|k_subpel_tma|k_subpel_exact|k_predict_u8 -s 30 -c 6`)\n")
    kn = h.index("Kernel Name")
    P("| metric | " + " | ".join(f"`{re.sub(r'[(].*', '', r[kn]).replace('void ', '')}`" for r in rr[2:]) + " |")
    P("|---|" + "---:|" * len(rr[2:]))
    for i, n in idx:
        P(f"| {n} ({units[i]}) | " + " | ".join(r[i] for r in rr[2:]) + " |")
    # SASS evidence
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "qsvc_b200", "libqsvc_b200.so")],
                          capture_output=True, text=True).stdout
    cnt = collections.Counter(re.findall(r"\b(UTMALDG\.2D|VABSDIFF4\.U8\.ACC|VABSDIFF4\.U8|VABSDIFF|REDUX\.SUM|SYNCS\.[A-Z.0-9]+)\b", sass))
    P("\n## SASS evidence (`cuobjdump -sass qsvc_b200/libqsvc_b200.so`)\n")
    for k, v in sorted(cnt.items()):
        P(f"* `{k}`: {v} sites")
out.close()
print(open(os.path.join(ROOT, "profiles", f"{tag}_summary.md")).read())
