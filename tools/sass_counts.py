"""Per-kernel SASS instruction counts of the built library (cuobjdump -sass): the Blackwell-native signature of the
FP64 path is DMMA.8x8x4 (no tcgen05 kind exists for f64), UTMALDG (TMA loads), USETMAXREG (warp-group register
re-allocation), with no UTCMMA / LDTM / STTM.  usage: python tools/sass_counts.py [out.json]"""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaussianprocesspathmodelling_b200", "libgpmap_b200.so")
OPS = ["DMMA", "DFMA", "DADD", "DMUL", "UTMALDG", "UTMASTG", "USETMAXREG", "UTCMMA", "LDTM", "STTM", "LDS", "STS", "LDG", "STG",
       "LDGSTS", "BAR", "SYNCS", "MUFU", "SHFL", "ATOMG", "RED", "FENCE", "MEMBAR", "LDL", "STL"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
out, cur = {}, None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = out.setdefault(name, {"instructions": 0})
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["instructions"] += 1
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o in ("LDS", "STS", "LDG", "STG", "BAR") and op.startswith(o)):
                cur[o] = cur.get(o, 0) + 1
                break
tot = {}
for k in out.values():
    for o, v in k.items():
        tot[o] = tot.get(o, 0) + v
res = {"library": os.path.relpath(LIB, ROOT), "arch": "sm_100a", "totals": tot, "kernels": dict(sorted(out.items()))}
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_counts.json")
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(tot))
