"""Executed warp instructions and stall samples per CUDA source line of one kernel of an
ncu report (needs -lineinfo at compile time and `--import-source on`).

  python profiles/by_line.py gpurun_out/x.ncu-rep 'k_mc_tile<(int)3' qsvc_b200/libqsvc_b200.so [top]

ncu's source page lists the kernel's SASS in program order with per-instruction counters;
nvdisasm -g lists the same instructions with `//## File "...", line N` markers.  Joined by
instruction index.
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, pat, so = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = next(i for i in secs if pat in rows[i][1])
end = next((j for j in secs if j > sec), len(rows))
h = rows[sec + 1]
ci, si, smp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
inst = [(r[si].strip(), int(r[ci] or 0), int(r[smp] or 0)) for r in rows[sec + 2:end] if len(r) > ci]
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", pat.split("<")[0])

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
best = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
    cur, line, seq = None, 0, {}
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur, line = m.group(1), 0
            seq[cur] = []
            continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and cur:
            seq[cur].append((line, m.group(1).strip()))
    for name, s in seq.items():
        if mangled_hint in name and len(s) == len(inst):
            # same length: check opcodes of the first instructions agree
            if all(a[1].split()[0].lstrip("@!P0123456789T ") [:3] == b[0].split()[0].lstrip("@!P0123456789T ")[:3]
                   for a, b in list(zip(s, inst))[:50] if a[1] and b[0]):
                best = s
if best is None:
    sys.exit("no function of matching length found for " + pat)
per = collections.Counter()
smp_per = collections.Counter()
tot = sum(i[1] for i in inst)
stot = sum(i[2] for i in inst)
for (line, _), (_, n, s) in zip(best, inst):
    per[line] += n
    smp_per[line] += s
print(f"{pat}: {tot} warp instructions, {stot} samples, {len(inst)} SASS instructions")
srcs = {}
for (f, l), n in per.most_common(top):
    if f not in srcs:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(so)), "**", f), recursive=True)
        srcs[f] = open(cands[0]).read().splitlines() if cands else []
    text = srcs[f][l - 1].strip() if 0 < l <= len(srcs[f]) else ""
    print(f"{n / tot * 100:5.1f}% inst {smp_per[(f, l)] / max(stot, 1) * 100:5.1f}% samp  {f}:{l:<4} {text[:100]}")
