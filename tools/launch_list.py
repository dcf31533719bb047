"""Print the second half (= last call) of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
d = [r for r in rows[hdr + 1:] if len(r) > vi and r[0].isdigit()]
tot = 0.0
for r in d[len(d) // 2:]:
    print("%-44s %10.1f us" % (r[ki][:44], float(r[vi].replace(",", "")) / 1e3)); tot += float(r[vi].replace(",", "")) / 1e3
print("total %.1f us" % tot)
