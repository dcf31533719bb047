"""Aggregate an `ncu --page source --csv` SASS export per CUDA source line (file-aware), using nvdisasm line info of
the cubin the kernel came from.  usage: ncu_by_line.py <sass.csv> <cubin> <mangled-name-substring> [min_pct]"""
import csv, os, re, subprocess, sys
from collections import defaultdict
sass_csv, cubin, sub = sys.argv[1:4]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], ("", 0), False
for t in txt:
    if t.startswith("\t.section") or t.startswith(".section"):
        on = (".text." in t) and (sub in t)
    if not on: continue
    m = re.search(r'//## File "(.*?)", line (\d+)', t)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", t): lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
h = rows[1]; ie = h.index("Instructions Executed"); ws = h.index("L1 Wavefronts Shared"); st = h.index("Warp Stall Sampling (All Samples)")
data = rows[2:]
print("sass rows", len(data), "nvdisasm instr", len(lines))
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
for i, r in enumerate(data):
    ln = lines[i] if i < len(lines) else ("", -1)
    a = agg[ln]; a[0] += float(r[ie] or 0); a[1] += float(r[ws] or 0); a[2] += float(r[st] or 0)
tot = [sum(a[k] for a in agg.values()) for k in range(3)]
srcs = {}
def src(f, ln):
    if f not in srcs:
        try: srcs[f] = open("/root/repo/gaussianprocesspathmodelling_b200/csrc/" + f).read().splitlines()
        except Exception: srcs[f] = []
    s = srcs[f]
    return s[ln - 1].strip()[:90] if 0 < ln <= len(s) else ""
for (f, ln) in sorted(agg):
    a = agg[(f, ln)]
    if a[0] > minpct / 100 * tot[0] or a[2] > minpct / 100 * tot[2]:
        print("%-12s %4d inst %5.1f%% wf %5.1f%% stall %5.1f%%  %s" % (f, ln, a[0] / tot[0] * 100, a[1] / max(tot[1], 1) * 100, a[2] / max(tot[2], 1) * 100, src(f, ln)))
