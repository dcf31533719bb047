"""Condense an ncu report (--set full) into one JSON line per launch: duration, DRAM bytes, pipe utilisation.
usage: ncu_summary.py report.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
def val(r, name, scale=None):
    if name not in h:
        return None
    i = h.index(name); v = r[i].replace(",", "")
    if v == "":
        return None
    x = float(v); u = units[i]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return x * mult
out = []
for r in rows[2:]:
    d = {"kernel": r[h.index("Kernel Name")].split("(")[0].replace("void ", ""),
         "grid": r[h.index("launch__grid_size")], "block": r[h.index("launch__block_size")],
         "duration_us": val(r, "gpu__time_duration.sum"),
         "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
         "dram_throughput_pct_of_ncu_peak": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
         "fp64_pipe_active_pct": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
         "dmma_pipe_active_pct": val(r, "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
         "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "l2_hit_rate_pct": val(r, "lts__t_sector_hit_rate.pct"),
         "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
         "registers_per_thread": val(r, "launch__registers_per_thread")}
    if d["duration_us"]:
        d["dram_gbs"] = ((d["dram_read_bytes"] or 0) + (d["dram_write_bytes"] or 0)) / d["duration_us"] / 1e3
    out.append(d)
txt = "[\n" + ",\n".join(" " + json.dumps(d) for d in out) + "\n]\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
print(txt)
