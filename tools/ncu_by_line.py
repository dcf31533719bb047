"""Aggregate an `ncu --page source --csv` SASS export per CUDA source line, using nvdisasm line info of the
cubin the kernel came from.  usage: ncu_by_line.py <sass.csv> <cubin> <mangled-name-substring>"""
import csv, re, subprocess, sys
from collections import defaultdict
sass_csv, cubin, sub = sys.argv[1:4]
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], 0, False
for t in txt:
    if t.startswith("\t.section") or t.startswith(".section"):
        on = (".text." in t) and (sub in t)
    if not on: continue
    m = re.search(r'//## File ".*?", line (\d+)', t)
    if m: cur = int(m.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", t): lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
h = rows[1]; ie = h.index("Instructions Executed"); ws = h.index("L1 Wavefronts Shared"); st = h.index("Warp Stall Sampling (All Samples)")
data = rows[2:]
print("sass rows", len(data), "nvdisasm instr", len(lines))
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
for i, r in enumerate(data):
    ln = lines[i] if i < len(lines) else -1
    a = agg[ln]; a[0] += float(r[ie] or 0); a[1] += float(r[ws] or 0); a[2] += float(r[st] or 0)
tot = [sum(a[k] for a in agg.values()) for k in range(3)]
src = open("/root/repo/gaussianprocesspathmodelling_b200/csrc/" + cubin.split("/")[-1].split(".")[0] + ".cu").read().splitlines()
for ln in sorted(agg):
    a = agg[ln]
    if a[0] > 0.004 * tot[0] or a[2] > 0.004 * tot[2]:
        print("%4d inst %5.1f%% wf %5.1f%% stall %5.1f%%  %s" % (ln, a[0] / tot[0] * 100, a[1] / max(tot[1], 1) * 100, a[2] / max(tot[2], 1) * 100, src[ln - 1].strip()[:100] if 0 < ln <= len(src) else ""))
