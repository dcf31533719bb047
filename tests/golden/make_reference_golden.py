"""Generate golden vectors from the UNMODIFIED reference (/root/reference/GPmap.py) for the code it
actually contains: CSV ingest, check_if_valid_trajectory, calc_distance, calc_mean_traj and one
k-means assignment step.  Run in the build container only (the reference does not travel to the GPU
box); the outputs are committed:  tests/golden/reference_testfile.csv, tests/golden/reference_golden.npz.

The reference executes its whole pipeline at import (GPmap.py:212,220) and needs matplotlib, which is
not installed: a 3-function stub package is put on sys.path, and the import runs in a scratch directory
holding the synthetic testfile.csv.
"""
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def synth_csv(path, n_traj=40, n=33, seed=7):
    rng = np.random.default_rng(seed)
    lines = []
    for k in range(n_traj):
        npts = n if k not in (3, 9) else n - 5              # two paths of the wrong length are dropped
        t = np.linspace(0, 1, npts)
        c = rng.uniform(-2e4, 2e4, 2)
        amp = rng.uniform(5e3, 2.5e4)
        ph = rng.uniform(0, 2 * np.pi)
        x = np.rint(c[0] + amp * np.cos(2 * np.pi * t * rng.uniform(0.3, 1.0) + ph) + 200 * rng.standard_normal(npts))
        y = np.rint(c[1] + amp * np.sin(2 * np.pi * t * rng.uniform(0.3, 1.0) + ph) + 200 * rng.standard_normal(npts))
        if k == 5:                                          # shrinking path: fails the validity filter
            x = np.rint(np.linspace(3e4, 10, npts)); y = np.rint(np.linspace(-3e4, -10, npts))
        lines.append(f"traj,id{k:03d},x,y")
        for i in range(npts):
            lines.append(f"{0.5 * i},p,{int(x[i])},{int(y[i])}")
        lines.append("###,,,")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def main():
    scratch = tempfile.mkdtemp(prefix="gpmap_ref_")
    os.makedirs(os.path.join(scratch, "matplotlib"))
    with open(os.path.join(scratch, "matplotlib", "__init__.py"), "w") as fh:
        fh.write("")
    with open(os.path.join(scratch, "matplotlib", "pyplot.py"), "w") as fh:
        fh.write("def axis(*a, **k): pass\ndef plot(*a, **k): pass\ndef show(*a, **k): pass\n")
    with open(os.path.join(scratch, "matplotlib", "cm.py"), "w") as fh:
        fh.write("")
    csv_path = os.path.join(HERE, "reference_testfile.csv")
    synth_csv(csv_path)
    os.chdir(scratch)
    os.symlink(csv_path, os.path.join(scratch, "testfile.csv"))
    sys.path.insert(0, scratch)
    sys.path.insert(0, REF)
    import warnings
    warnings.simplefilter("ignore")
    random.seed(12345)
    import GPmap as ref                                      # runs readcsvfile(10) + kmeansclustering(3)

    # full ingest (the import stopped at 10 paths)
    ref.trajs = ref.trajectories()
    ref.readcsvfile(0)
    T = ref.trajs
    keys = list(T.pathdict.keys())
    xs = np.stack([T.pathdict[k].xs for k in keys]); ys = np.stack([T.pathdict[k].ys for k in keys])
    ts = np.stack([T.pathdict[k].timestamp for k in keys])
    P = len(keys)
    dist = np.array([[T.calc_distance(T.pathdict[a], T.pathdict[b]) for b in keys] for a in keys])
    # validity sums via the reference's own double loop (threshold sweep gives the sum's sign only, so
    # re-run the loop body here exactly as GPmap.py:169-172 does)
    def ref_sum(tr):
        s = 0
        for i in range(len(tr.xs)):
            for j in range(i + 1, len(tr.xs)):
                s += abs(tr.xs[j]) - abs(tr.xs[i])
                s += abs(tr.ys[j]) - abs(tr.ys[i])
        return s
    sums = np.array([ref_sum(T.pathdict[k]) for k in keys])
    valid_1000 = np.array([ref.check_if_valid_trajectory(T.pathdict[k], 1000) for k in keys])
    # centroid means of three fixed groups and the distances of every path to them (one Lloyd step)
    third = max(1, P // 3)
    groups = [keys[0:third], keys[third:2 * third], keys[2 * third:]]
    cents = [T.calc_mean_traj(g) for g in groups]
    cx = np.stack([c.xs for c in cents]); cy = np.stack([c.ys for c in cents]); ct = np.stack([c.timestamp for c in cents])
    d2c = np.array([[T.calc_distance(c, T.pathdict[k]) for c in cents] for k in keys])
    assign = np.array([int(np.argmin(r)) for r in d2c])      # strict '<' first-minimum == argmin
    np.savez(os.path.join(HERE, "reference_golden.npz"), keys=np.array(keys), xs=xs, ys=ys, ts=ts, dist=dist,
             sums=sums, valid_1000=valid_1000, group_sizes=np.array([len(g) for g in groups]),
             cx=cx, cy=cy, ct=ct, d2c=d2c, assign=assign)
    print("kept", P, "paths:", keys)
    print("sums", sums)


if __name__ == "__main__":
    main()
