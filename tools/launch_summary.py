"""Per-kernel shares of the LAST call in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_summary.py <launches.csv> <n_calls_in_list> "<what>" [out.json]
The list holds n identical calls back to back (warm-ups + the measured one); the launches of the last one are summed
per kernel.  ncu durations are cold-cache and serialised: compare shares, not absolutes."""
import csv, json, re, sys
from collections import OrderedDict
path, ncalls, what = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
d = [r for r in rows[hdr + 1:] if len(r) > vi and r[0].isdigit()]
ours = [r for r in d if "gpm::" in r[ki]]          # the library's kernels (torch fill / check kernels excluded)
last = ours[len(ours) - len(ours) // ncalls:]
def name(k):
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"\(.*$", "", k)
    return re.sub(r"\(int\)", "", k)
agg = OrderedDict(); tot = 0.0
for r in last:
    us = float(r[vi].replace(",", "")) / (1e3 if r[ui].startswith("ns") else 1.0)
    a = agg.setdefault(name(r[ki]), [0, 0.0]); a[0] += 1; a[1] += us; tot += us
out = {"what": what, "total_us": round(tot, 1), "launches": len(last),
       "kernels": {k: {"launches": v[0], "us": round(v[1], 1), "share": round(v[1] / tot, 4)}
                   for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
txt = json.dumps(out, indent=1)
print(txt)
if len(sys.argv) > 4:
    open(sys.argv[4], "w").write(txt + "\n")
