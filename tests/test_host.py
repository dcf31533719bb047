"""Host-side mirror of the reference's data model (GPmap.py:12-34,165-204), checked against golden
vectors generated from the reference itself (tests/golden/make_reference_golden.py)."""
import os

import numpy as np

from gaussianprocesspathmodelling_b200 import GPmap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(ROOT, "tests", "golden", "reference_testfile.csv")


def test_import_has_no_side_effects():
    assert isinstance(GPmap.trajs, GPmap.trajectories) and GPmap.trajs.pathdict == {}


def test_trajectory_conventions():
    t = GPmap.trajectory()
    assert t.xs.dtype == np.float64 and t.xs.shape == (0,)
    t.add_point(0.5, 3, -4)
    t.add_point(1.0, 5, 6)
    assert t.xs.tolist() == [3.0, 5.0] and t.ys.tolist() == [-4.0, 6.0] and t.timestamp.tolist() == [0.5, 1.0]
    assert t.gp_inputs().shape == (2, 2) and t.gp_inputs(use_time=True).shape == (2, 3)
    assert t.get_trajectory().tolist() == [[0.5, 3.0, -4.0], [1.0, 5.0, 6.0]]


def test_readcsvfile_matches_reference(ref_golden):
    dest = GPmap.readcsvfile(0, filename=CSV, target=GPmap.trajectories())
    assert list(dest.pathdict.keys()) == ref_golden["keys"].tolist()
    for i, k in enumerate(dest.pathdict):
        assert np.array_equal(dest.pathdict[k].xs, ref_golden["xs"][i])
        assert np.array_equal(dest.pathdict[k].ys, ref_golden["ys"][i])
        assert np.array_equal(dest.pathdict[k].timestamp, ref_golden["ts"][i])
    few = GPmap.readcsvfile(4, filename=CSV, target=GPmap.trajectories())
    assert list(few.pathdict.keys()) == ref_golden["keys"].tolist()[:4]


def test_validity_filter_matches_reference(ref_golden):
    T = GPmap.readcsvfile(0, filename=CSV, target=GPmap.trajectories())
    for i, k in enumerate(T.pathdict):
        s = ref_golden["sums"][i]
        for thr in (1, 1000, s, s + 1, s - 1):
            assert GPmap.check_if_valid_trajectory(T.pathdict[k], thr) == (not (s < thr))
    shrinking = GPmap.trajectory()
    for i in range(10):
        shrinking.add_point(i, 100 - 10 * i, 0)
    assert GPmap.check_if_valid_trajectory(shrinking, 1) is False


def test_distance_and_mean_match_reference(ref_golden):
    T = GPmap.readcsvfile(0, filename=CSV, target=GPmap.trajectories())
    keys = list(T.pathdict)
    for a in range(len(keys)):
        for b in range(len(keys)):
            assert T.calc_distance(T.pathdict[keys[a]], T.pathdict[keys[b]]) == ref_golden["dist"][a, b]
    o = 0
    for c, n in enumerate(ref_golden["group_sizes"]):
        m = T.calc_mean_traj(keys[o:o + n])
        assert np.array_equal(m.xs, ref_golden["cx"][c]) and np.array_equal(m.ys, ref_golden["cy"][c])
        assert np.array_equal(m.timestamp, ref_golden["ct"][c])
        for p, k in enumerate(keys):
            # fractional centroid coordinates: the reference's np.linalg.norm goes through BLAS ddot,
            # whose FMA use is CPU-dependent, so equality holds only to an ulp or two
            assert abs(T.calc_distance(m, T.pathdict[k]) - ref_golden["d2c"][p, c]) <= 1e-15 * ref_golden["d2c"][p, c]
        o += n


def test_packed_and_theta():
    T = GPmap.readcsvfile(0, filename=CSV, target=GPmap.trajectories())
    Xb, keys = T.packed(use_time=True)
    assert Xb.shape == (len(keys), 33, 3) and Xb.dtype == np.float64
    assert GPmap.make_theta(8000.0, 1.0, 0.01, 3).tolist() == [8000.0, 8000.0, 8000.0, 1.0, 0.01]
    assert GPmap.make_theta([1.0, 2.0], 3.0, 4.0, 2).tolist() == [1.0, 2.0, 3.0, 4.0]


def test_lloyd_restatement_matches_the_reference_run():
    """oracle.gp_ref.lloyd against a full run of the reference's own kmeansclustering (fixed initial centroids,
    tests/golden/make_reference_kmeans_golden.py): same iteration count, assignment and bit-identical centroids."""
    from oracle import gp_ref
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_kmeans_golden.npz"))
    for tag in ("a", "b"):
        keys = g[f"{tag}_keys"].tolist()
        init = [keys.index(k) for k in g[f"{tag}_init"].tolist()]
        assign, cents, iters = gp_ref.lloyd(g[f"{tag}_xs"], g[f"{tag}_ys"], g[f"{tag}_ts"], init)
        assert iters == int(g[f"{tag}_iters"])
        assert np.array_equal(assign, g[f"{tag}_assign"])
        assert np.array_equal(cents, g[f"{tag}_cents"])
