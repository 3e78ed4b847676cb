"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries in profiles/.

  python profiles/summarize.py r1
reads
  gpurun_out/launches_<tag>.csv   ncu --metrics gpu__time_duration.sum          (launch list of one step)
  gpurun_out/metrics_<tag>.csv    ncu --metrics <time, dram bytes, instructions, issue, occupancy> (same step)
  gpurun_out/top_<tag>.ncu-rep    ncu --set full of the level-1 launches of the top kernels
writes
  profiles/<tag>_summary.md, profiles/<tag>_launches.csv (copy), profiles/<tag>_ncu.json
    (dram bytes per step per bench.py kernel class: the `roofline.traffic` source)
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w")
P = lambda *a: print(*a, file=out)
SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "s": 1e3}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

# bench.py kernel classes (csrc/kernels.cuh KC_*)
CLASS = [("residue", r"k_mc_march|k_ll_residue|k_residue"), ("search_exact", r"k_subpel_strip|k_subpel_exact|k_strip2|k_level1_tile"),
         ("search", r"k_subpel_tma|k_subpel_fast|k_search"), ("predict", r"k_predict|k_tail_state|k_clip"),
         ("dwt_rows", r"k_dwt_rows|k_dwt0_u8|k_dwt_snap|k_ll1_store_u8"), ("dwt_cols", r"k_dwt_cols|k_syn_snap"), ("update", r"k_update"), ("image", r".*")]


def kname(s):
    return re.sub(r"\(.*", "", s).replace("void ", "")


def klass(name):
    for c, pat in CLASS:
        if re.match(pat, name):
            return c
    return "image"


def long_csv(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    idx = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    for r in rows[hi + 1:]:
        if len(r) > idx["Metric Value"]:
            yield (r[idx["ID"]], kname(r[idx["Kernel Name"]]), r[idx["Metric Name"]], r[idx["Metric Unit"]],
                   float(r[idx["Metric Value"]].replace(",", "") or 0))


# ---- launch list: per-kernel share of one analysis step (cold-cache, serialised)
src = os.path.join(GO, f"launches_{tag}.csv")
agg = collections.OrderedDict()
for _id, name, metric, unit, v in long_csv(src):
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v * SCALE.get(unit, 1e-6)
shutil.copy(src, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
tot = sum(v[1] for v in agg.values())
n_launch = sum(v[0] for v in agg.values())
P(f"# ncu summaries, round {tag}\n")
P("Workload: `python profiles/run_step.py cfg3 2` = bench.py's cfg3 (1080p, 129 frames, GOP 32, block 16,")
P("search 16, quarter-pel), second (warm) analysis step.  Launch list:")
P(f"`ncu --metrics gpu__time_duration.sum --clock-control none -s {n_launch} -c {n_launch}` (times are cold-cache and")
P("serialised: compare shares, not absolutes).\n")
P(f"One step = {n_launch} kernel launches, {tot:.1f} ms summed under ncu.\n")
P("| kernel | class (bench.py) | launches | ms (ncu) | share |")
P("|---|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    P(f"| `{k}` | {klass(k)} | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")
cls_ms = collections.Counter()
for k, v in agg.items():
    cls_ms[klass(k)] += v[1]
P("\nPer class: " + ", ".join(f"{c} {ms:.2f} ms ({100 * ms / tot:.0f}%)" for c, ms in cls_ms.most_common()) + "\n")

# ---- per-launch metrics of the same step: DRAM traffic, instructions, issue utilisation
mpath = os.path.join(GO, f"metrics_{tag}.csv")
traffic_cls = collections.Counter()
if os.path.exists(mpath):
    per = collections.defaultdict(dict)
    names = {}
    for _id, name, metric, unit, v in long_csv(mpath):
        names[_id] = name
        if metric.startswith("dram__bytes"):
            v *= BYTES.get(unit, 1.0)
        elif metric == "gpu__time_duration.sum":
            v *= SCALE.get(unit, 1e-6)
        per[_id][metric] = v
    k = collections.OrderedDict()
    for _id, m in per.items():
        a = k.setdefault(names[_id], collections.Counter())
        t = m.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["ms"] += t
        a["rd"] += m.get("dram__bytes_read.sum", 0.0)
        a["wr"] += m.get("dram__bytes_write.sum", 0.0)
        a["inst"] += m.get("smsp__inst_executed.sum", 0.0)
        a["issue_w"] += t * m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
        a["occ_w"] += t * m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
        a["regs"] = max(a["regs"], m.get("launch__registers_per_thread", 0.0))
        traffic_cls[klass(names[_id])] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    P("## Per-kernel counters of the same step (`ncu --metrics ...`, summed over the step's launches)\n")
    P("| kernel | launches | ms | DRAM read GB | DRAM write GB | DRAM GB/s | warp instr (G) | issue slots busy % | achieved occupancy % | regs |")
    P("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for name, a in sorted(k.items(), key=lambda kv: -kv[1]["ms"]):
        ms = max(a["ms"], 1e-9)
        P(f"| `{name}` | {int(a['n'])} | {a['ms']:.2f} | {a['rd'] / 1e9:.2f} | {a['wr'] / 1e9:.2f} | "
          f"{(a['rd'] + a['wr']) / 1e9 / (ms * 1e-3):.0f} | {a['inst'] / 1e9:.2f} | {a['issue_w'] / ms:.0f} | "
          f"{a['occ_w'] / ms:.0f} | {int(a['regs'])} |")
    P("")
json.dump({"workload": "cfg3", "source": f"gpurun_out/metrics_{tag}.csv (ncu, one warm step)",
           "dram_bytes_per_step": {c: v for c, v in traffic_cls.items()}},
          open(os.path.join(ROOT, "profiles", f"{tag}_ncu.json"), "w"), indent=1)

# ---- full captures of the top kernels
import glob
reps = sorted(glob.glob(os.path.join(GO, f"top_{tag}*.ncu-rep")))
cols_all, col_units, h, units = [], [], None, None
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr_ = list(csv.reader(raw.splitlines()))
    if len(rr_) < 3:
        continue
    if h is None:
        h, units = rr_[0], rr_[1]
        cols_all += rr_[2:]
        col_units += [rr_[1]] * len(rr_[2:])
    else:  # align the columns of further reports to the first header (units differ per report)
        pos = {n: i for i, n in enumerate(rr_[0])}
        for r in rr_[2:]:
            cols_all.append([r[pos[n]] if n in pos and pos[n] < len(r) else "" for n in h])
            col_units.append([rr_[1][pos[n]] if n in pos else "" for n in h])
if h is not None:
    rr = [h, units] + cols_all
    want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"),
            ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
            ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("launch__registers_per_thread", "registers/thread"),
            ("smsp__inst_executed.sum", "warp instructions"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
            ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard"),
            ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait"),
            ("smsp__pcsamp_warps_issue_stalled_barrier", "stall samples: barrier"),
            ("smsp__pcsamp_warps_issue_stalled_no_instructions", "stall samples: no instructions"),
            ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall samples: math pipe throttle"),
            ("smsp__pcsamp_sample_count", "samples"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block")]
    idx = [(h.index(m), n) for m, n in want if m in h]
    ki = h.index("Kernel Name")
    cols = rr[2:]
    P("## `ncu --set full --clock-control none --import-source on` of the level-1 launches of the top kernels\n")
    P("(first temporal level of the step: 64 frame pairs; `profiles/capture.sh`; k_mc_march from a second capture\n"
      "with `-k regex:k_mc_march -c 2`)\n")
    P("| metric | " + " | ".join(f"`{kname(r[ki])}`" for r in cols) + " |")
    P("|---|" + "---:|" * len(cols))
    for i, n in idx:
        P(f"| {n} | " + " | ".join(f"{r[i]} {u[i]}".strip() for r, u in zip(cols, col_units)) + " |")
    P("")

# ---- SASS evidence
so = os.path.join(ROOT, "qsvc_b200", "libqsvc_b200.so")
if os.path.exists(so):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    P("## SASS evidence (`cuobjdump -sass qsvc_b200/libqsvc_b200.so`)\n")
    for pat in ("REDUX.SUM", "SYNCS.ARRIVE.TRANS64", "SYNCS.EXCH.64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "UTMALDG.2D",
                "VABSDIFF4.U8.ACC", "VABSDIFF "):
        P(f"* `{pat.strip()}`: {sass.count(pat)} sites")
out.close()
print(open(os.path.join(ROOT, "profiles", f"{tag}_summary.md")).read())
