"""GPU parity tests: the CUDA path (through the C-ABI, via the GPmap drop-in module) against the
numpy/scipy oracle and the committed golden vectors.  Tolerances are BASELINE.json's: 1e-8 on the
posterior mean, 1e-6 on the variance, both norm-wise relative (max|d| / max|ref|); LML 1e-10 relative.
"""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import gp_ref                                                      # noqa: E402
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl  # noqa: E402

MEAN_TOL, VAR_TOL, LML_TOL = 1e-8, 1e-6, 1e-10


def nrm(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_cov(X, th, lower=False):
    lib = _native.load(); h = _native.handle(0)
    N, D = X.shape
    ld = (N + 15) // 16 * 16
    K = torch.full((N, ld), float("nan"), dtype=torch.float64, device="cuda")
    Xd = dev(X)
    rc = lib.gpm_cov(h, C.c_void_p(Xd.data_ptr()), N, D, _native.theta_array(th), C.c_void_p(K.data_ptr()), ld,
                     1 if lower else 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _native.check(rc, "gpm_cov")
    torch.cuda.synchronize()
    return K[:, :N].cpu().numpy()


def gpu_potrf(Knp):
    lib = _native.load(); h = _native.handle(0)
    N = Knp.shape[0]
    ld = (N + 15) // 16 * 16
    K = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
    K[:, :N] = dev(Knp)
    ws = torch.empty(int(lib.gpm_potrf_workspace_bytes(N)) // 8, dtype=torch.float64, device="cuda")
    info = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    rc = lib.gpm_potrf(h, C.c_void_p(K.data_ptr()), N, ld, C.c_void_p(ws.data_ptr()), C.c_void_p(info.data_ptr()),
                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _native.check(rc, "gpm_potrf")
    torch.cuda.synchronize()
    return np.tril(K[:, :N].cpu().numpy()), ws, int(info.item())


@pytest.mark.parametrize("N,D", [(1, 2), (33, 2), (33, 3), (64, 2), (65, 3), (200, 2), (511, 3), (1000, 2)])
def test_cov_matches_oracle(N, D):
    X, _, th = wl.single_path(max(N, 4), seed=100 + N, D=D)
    X = X[:N]
    K = gpu_cov(X, th)
    Ko = gp_ref.cov(X, th)
    assert np.abs(K - Ko).max() < 4e-15          # entries are <= 1.01; exp() implementations differ by < 1 ulp
    assert np.array_equal(K, K.T)                # mirrored tiles are bitwise symmetric
    Kl = gpu_cov(X, th, lower=True)
    assert np.array_equal(np.tril(Kl), np.tril(K))


@pytest.mark.parametrize("N", [1, 33, 128, 129, 200, 256, 300, 511, 1000, 1537])
def test_potrf_matches_scipy(N):
    X, _, th = wl.single_path(max(N, 4), seed=200 + N, D=2)
    Ko = gp_ref.cov(X[:N], th)
    L, ws, info = gpu_potrf(Ko)
    assert info == 0
    Lo = scipy.linalg.cholesky(Ko, lower=True)
    assert np.abs(L @ L.T - Ko).max() / np.abs(Ko).max() < 1e-14
    assert nrm(L, Lo) < 1e-11
    # inverted diagonal blocks
    nblk = (N + 127) // 128
    inv = ws.cpu().numpy().reshape(nblk, 128, 128)
    for k in range(nblk):
        nv = min(128, N - 128 * k)
        blk = Lo[128 * k:128 * k + nv, 128 * k:128 * k + nv]
        assert np.abs(inv[k][:nv, :nv] @ blk - np.eye(nv)).max() < 1e-11
        assert np.array_equal(inv[k][nv:, nv:], np.eye(128 - nv))


def test_potrf_lookahead_equals_plain():
    X, _, th = wl.single_path(900, seed=5, D=2)
    Ko = gp_ref.cov(X, th)
    L1, _, _ = gpu_potrf(Ko)
    with _native.option("no_lookahead", 1):
        L2, _, _ = gpu_potrf(Ko)
    assert np.array_equal(L1, L2)                # same kernels, same order of arithmetic


def test_potrf_split_column_lookahead_equals_plain():
    # 24 <= blocks < 36: the look-ahead updates the next diagonal tile first and the rest of its column on the
    # caller's stream (csrc/potrf.cu); same kernels on the same tiles, so the factor is bitwise the unsplit one
    N = 3100                                     # 25 blocks, ragged last block
    X, _, th = wl.single_path(N, seed=6, D=2)
    Ko = gp_ref.cov(X, th)
    L1, _, info = gpu_potrf(Ko)
    assert info == 0
    with _native.option("no_split_column", 1):
        L2, _, _ = gpu_potrf(Ko)
    with _native.option("no_lookahead", 1):
        L3, _, _ = gpu_potrf(Ko)
    assert np.array_equal(L1, L2) and np.array_equal(L1, L3)
    assert np.abs(L1 @ L1.T - Ko).max() / np.abs(Ko).max() < 1e-14
    # fused forward substitution rides along the split schedule too
    Y = np.random.default_rng(7).standard_normal((N, 2))
    m = GPmap.fit_gp(X, Y, theta=th)
    with _native.option("no_split_column", 1):
        m2 = GPmap.fit_gp(X, Y, theta=th)
    assert torch.equal(m.alpha, m2.alpha) and torch.equal(m.lml_dev, m2.lml_dev)


def test_potrf_reports_non_positive_pivot():
    A = np.eye(300); A[150, 150] = -1.0
    _, _, info = gpu_potrf(A)
    assert info == 151
    B = np.eye(40); B[3, 3] = 0.0
    assert gpu_potrf(B)[2] == 4
    with pytest.raises(np.linalg.LinAlgError):
        GPmap.fit_gp(np.zeros((5, 2)), np.ones(5), lengthscale=1.0, noise_var=0.0)   # singular: all points equal


@pytest.mark.parametrize("tag,D", [("n33d2", 2), ("n33d3", 3)])
def test_fit_predict_golden_n33(golden, tag, D):
    X, Y, th = golden[f"{tag}_X"], golden[f"{tag}_Y"], golden[f"{tag}_theta"]
    m = GPmap.fit_gp(X, Y, theta=th)
    assert nrm(np.tril(m.L.cpu().numpy()), golden[f"{tag}_L"]) < 1e-12
    assert nrm(m.alpha.cpu().numpy(), golden[f"{tag}_alpha"]) < MEAN_TOL
    assert np.abs(m.lml - golden[f"{tag}_lml"]).max() < LML_TOL * np.abs(golden[f"{tag}_lml"]).max()
    mu, var = m.predict_grid(wl.BOX, (9, 7), t=7.5 if D == 3 else None)
    assert mu.shape == (7, 9, 2) and var.shape == (7, 9)
    assert nrm(mu.cpu().numpy(), golden[f"{tag}_mu"]) < MEAN_TOL
    assert nrm(var.cpu().numpy(), golden[f"{tag}_var"]) < VAR_TOL


def test_config1_golden(golden):
    X, Y, th = wl.single_path(200, seed=1, D=2, R=2)
    m = GPmap.fit_gp(X, Y, lengthscale=8000.0, signal_var=1.0, noise_var=1e-2)
    assert nrm(m.alpha.cpu().numpy(), golden["cfg1_alpha"]) < MEAN_TOL
    assert np.abs(m.lml - golden["cfg1_lml"]).max() < LML_TOL * np.abs(golden["cfg1_lml"]).max()
    mu, var = m.predict_grid(wl.BOX, (100, 100))
    assert nrm(mu.cpu().numpy(), golden["cfg1_mu"]) < MEAN_TOL
    assert nrm(var.cpu().numpy(), golden["cfg1_var"]) < VAR_TOL
    # and against scikit-learn's independent numbers
    assert nrm(mu.cpu().numpy()[:, :, 0], golden["cfg1_sk_mu0"]) < MEAN_TOL
    assert nrm(var.cpu().numpy(), golden["cfg1_sk_var"]) < VAR_TOL
    # mean-only path and predictive (noisy) variance
    mu2 = m.predict_grid(wl.BOX, (100, 100), return_var=False)       # separate mean kernel (other summation order)
    assert float((mu2 - mu).abs().max()) <= 1e-12 * float(mu.abs().max())
    _, var_y = m.predict_grid(wl.BOX, (100, 100), include_noise=True)
    assert torch.allclose(var_y, var + 1e-2, rtol=0, atol=1e-15)


def test_ragged_n300_scattered_queries_golden(golden):
    X, Y, th = wl.single_path(300, seed=12, D=3, R=1)
    m = GPmap.fit_gp(X, Y, theta=th)
    assert nrm(m.alpha.cpu().numpy(), golden["n300_alpha"]) < MEAN_TOL
    assert abs(m.lml[0] - golden["n300_lml"][0]) < LML_TOL * abs(golden["n300_lml"][0])
    mu, var = m.predict(golden["n300_Xs"])
    assert nrm(mu.cpu().numpy(), golden["n300_mu"]) < MEAN_TOL
    assert nrm(var.cpu().numpy(), golden["n300_var"]) < VAR_TOL


@pytest.mark.parametrize("N,D,R,M", [(1, 2, 1, 5), (2, 3, 2, 130), (127, 2, 1, 1), (128, 2, 3, 128), (129, 3, 2, 300),
                                      (700, 2, 8, 1000), (1100, 3, 1, 257)])
def test_fit_predict_vs_oracle_ragged(N, D, R, M):
    rng = np.random.default_rng(N * 7 + D)
    X, Y2, th = wl.single_path(max(N, 4), seed=300 + N, D=D, R=2)
    X = X[:N]
    Y = rng.standard_normal((N, R))
    Xs = np.column_stack([rng.uniform(-5e4, 5e4, M), rng.uniform(-5e4, 5e4, M)] + ([rng.uniform(0, 50, M)] if D == 3 else []))
    mo = gp_ref.fit(X, Y, th)
    mu_o, var_o = gp_ref.predict(mo, Xs)
    m = GPmap.fit_gp(X, Y, theta=th)
    assert nrm(m.alpha.cpu().numpy(), mo["alpha"]) < MEAN_TOL
    assert np.abs(m.lml - mo["lml"]).max() < LML_TOL * np.abs(mo["lml"]).max()
    mu, var = m.predict(Xs)
    assert nrm(mu.cpu().numpy(), mu_o) < MEAN_TOL
    assert nrm(var.cpu().numpy(), var_o) < VAR_TOL


def test_grid_sharding_concatenates_to_the_full_grid():
    X, Y, th = wl.single_path(260, seed=9, D=2, R=2)
    m = GPmap.fit_gp(X, Y, theta=th)
    mu, var = m.predict_grid(wl.BOX, (37, 23))
    from gaussianprocesspathmodelling_b200.dist import shard_range
    parts = [m.predict_grid(wl.BOX, (37, 23), points=shard_range(37 * 23, r, 3)) for r in range(3)]
    assert torch.equal(torch.cat([p[0] for p in parts]), mu.view(-1, 2))
    assert torch.allclose(torch.cat([p[1] for p in parts]), var.view(-1), rtol=0, atol=1e-13)


def test_batched_golden_and_oracle(golden):
    Xb, Yb, th = wl.batched_paths(6, 130, seed=3, D=3, R=2)
    alpha, lml = GPmap.fit_gp_batched(Xb, Yb, theta=th)
    assert nrm(alpha.cpu().numpy(), golden["b6_alpha"]) < MEAN_TOL
    assert np.abs(lml.cpu().numpy() - golden["b6_lml"]).max() < LML_TOL * np.abs(golden["b6_lml"]).max()
    Xb, Yb, th = wl.batched_paths(5, 512, seed=3, D=3, R=2, first=40)
    alpha, lml = GPmap.fit_gp_batched(Xb, Yb, theta=th)
    a_o, l_o = gp_ref.fit_batched(Xb, Yb, th)
    assert nrm(alpha.cpu().numpy(), a_o) < MEAN_TOL
    assert np.abs(lml.cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max()
    # a batched fit equals the single fits
    one = GPmap.fit_gp(Xb[2], Yb[2], theta=th)
    assert nrm(alpha[2].cpu().numpy(), one.alpha.cpu().numpy()) < 1e-12

def test_batched_fused_forward_matches_separate_solve():
    """Tiled batched pipeline (option no_path_fused): the forward substitution fused into the factorisation
    (default for N <= 2048) against the solve kernel running both passes (option no_fused_fwd), for odd / full
    right-hand-side counts."""
    for (B, N, D, R) in ((5, 300, 2, 3), (3, 130, 3, 1), (2, 512, 2, 8)):
        Xb, Yb, th = wl.batched_paths(B, N, seed=4, D=D, R=2)
        rng = np.random.default_rng(R)
        Yb = np.ascontiguousarray(np.concatenate([Yb, Yb.std() * rng.standard_normal((B, N, 6))], axis=2)[:, :, :R])
        with _native.option("no_path_fused", 1):
            a1, l1 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
            with _native.option("no_fused_fwd", 1):
                a0, l0 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
        a_o, l_o = gp_ref.fit_batched(Xb, Yb, th)
        assert nrm(a1.cpu().numpy(), a0.cpu().numpy()) < 1e-12
        assert nrm(a1.cpu().numpy(), a_o) < MEAN_TOL
        assert np.abs(l1.cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max()
        assert np.abs(l1.cpu().numpy() - l0.cpu().numpy()).max() < 1e-12 * np.abs(l_o).max()


def test_batched_scratch_factor_store_is_bitwise_neutral():
    """Tiled batched fits store only the diagonal 8 x 8 tiles of every L_kk (L is scratch there: the log-determinant reads
    the diagonal, the backward pass inv(L_kk)) and do not generate the upper 64 x 64 quadrant of the diagonal covariance
    blocks; the result must be bitwise the one with the whole factor / whole tiles stored (option no_scratch_factor),
    ragged last block and many right-hand sides included, on a NaN-filled workspace."""
    for (B, N, D, R) in ((70, 300, 2, 3), (64, 512, 3, 1), (66, 130, 2, 8)):     # batch >= 64: the 8-warp batched potf2
        Xb, Yb, th = wl.batched_paths(B, N, seed=6, D=D, R=2)
        rng = np.random.default_rng(R)
        Yb = np.ascontiguousarray(np.concatenate([Yb, Yb.std() * rng.standard_normal((B, N, 6))], axis=2)[:, :, :R])
        npad = (N + 127) // 128 * 128
        ws = torch.empty(B * (npad * npad + npad * 128 + 8 + npad * 8), dtype=torch.float64, device="cuda")
        with _native.option("no_path_fused", 1):
            ws.fill_(float("nan"))            # whatever is not generated / stored stays NaN: a stray reader would show
            a1, l1 = GPmap.fit_gp_batched(Xb, Yb, theta=th, workspace=ws)
            a1, l1 = a1.clone(), l1.clone()
            with _native.option("no_scratch_factor", 1):
                ws.fill_(float("nan"))
                a0, l0 = GPmap.fit_gp_batched(Xb, Yb, theta=th, workspace=ws)
        assert torch.equal(a1, a0) and torch.equal(l1, l0)
        a_o, l_o = gp_ref.fit_batched(Xb[:3], Yb[:3], th)
        assert nrm(a1[:3].cpu().numpy(), a_o) < MEAN_TOL
        assert np.abs(l1[:3].cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max()


def test_lml_sweep_vs_oracle():
    X, Y, _ = wl.single_path(180, seed=5, D=2, R=2)
    ths = wl.sweep_thetas(D=2)[::9]
    got = GPmap.lml_sweep(X, Y, ths)
    want = gp_ref.lml_sweep(X, Y, ths)
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()


def test_kmeans_assign_matches_reference(ref_golden):
    g = ref_golden
    T = GPmap.trajectories()
    for i, k in enumerate(g["keys"].tolist()):
        t = GPmap.trajectory(); t.xs, t.ys, t.timestamp = g["xs"][i], g["ys"][i], g["ts"][i]
        T.add_trajectory(k, t)
    cents = []
    for c in range(len(g["group_sizes"])):
        t = GPmap.trajectory(); t.xs, t.ys, t.timestamp = g["cx"][c], g["cy"][c], g["ct"][c]
        cents.append(t)
    assign, dist = T._assign(list(T.pathdict), cents)
    assert np.array_equal(assign, g["assign"])
    assert np.abs(dist - g["d2c"]).max() <= 1e-15 * g["d2c"].max()
    clusters = T.kmeansclustering(3, seed=1)
    assert sorted(k for v in clusters.values() for k in v) == sorted(T.pathdict)


def test_kmeans_device_lloyd_matches_the_reference_run():
    """The device-resident Lloyd loop (assign + segmented mean + convergence kernels, paths uploaded once) against a
    full run of the reference's own kmeansclustering from the same initial centroids
    (tests/golden/reference_kmeans_golden.npz): same iteration count and assignment, bit-identical centroids."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kmeans_golden.npz"))
    for tag in ("a", "b"):
        keys = g[f"{tag}_keys"].tolist()
        T = GPmap.trajectories()
        for i, k in enumerate(keys):
            t = GPmap.trajectory(); t.xs, t.ys, t.timestamp = g[f"{tag}_xs"][i], g[f"{tag}_ys"][i], g[f"{tag}_ts"][i]
            T.add_trajectory(k, t)
        init = g[f"{tag}_init"].tolist()
        clusters = T.kmeansclustering(len(init), init=init)
        assert T.kmeans_iterations == int(g[f"{tag}_iters"])
        names = list(clusters)
        got = np.array([next(c for c, nm in enumerate(names) if key in clusters[nm]) for key in keys])
        assert np.array_equal(got, g[f"{tag}_assign"])
        cents = np.stack([np.stack([T.centroids[nm].xs for nm in names]), np.stack([T.centroids[nm].ys for nm in names]),
                          np.stack([T.centroids[nm].timestamp for nm in names])])
        assert np.array_equal(cents, g[f"{tag}_cents"])
        for nm in names:                                       # members listed in path order, as the reference appends them
            assert clusters[nm] == [k for k in keys if k in clusters[nm]]
    # an empty cluster keeps its centroid instead of dividing by zero (GPmap.py:111): two identical initial centroids
    T = GPmap.trajectories()
    for i, k in enumerate(g["a_keys"].tolist()):
        t = GPmap.trajectory(); t.xs, t.ys, t.timestamp = g["a_xs"][i], g["a_ys"][i], g["a_ts"][i]
        T.add_trajectory(k, t)
    dup = GPmap.trajectory(); dup.xs, dup.ys, dup.timestamp = g["a_xs"][0].copy(), g["a_ys"][0].copy(), g["a_ts"][0].copy()
    T.add_trajectory("dup", dup)
    clusters = T.kmeansclustering(3, init=[g["a_keys"][0], "dup", g["a_keys"][7]])
    assert sorted(k for v in clusters.values() for k in v) == sorted(T.pathdict)
    assert T.kmeans_iterations >= 1 and all(np.isfinite(c.xs).all() for c in T.centroids.values())


def test_kmeans_many_clusters_against_the_restated_loop():
    """More clusters than one assignment pass takes (eight per pass), sliced centroid sums (k = 20: seven slices of 16
    row entries; k = 150: one slice per centroid) and clusters spanning several 2048-path chunks, against the oracle's
    restatement of the reference's loops (pinned by the reference-golden tests): same assignment, bit-identical
    centroids after the same number of iterations."""
    for (P, k, iters) in ((5000, 20, 2), (600, 150, 2)):
        xs, ys, ts = wl.trajectory_families(P, min(k, 12), 33, seed=31)
        T = GPmap.trajectories()
        for i in range(P):
            t = GPmap.trajectory(); t.xs, t.ys, t.timestamp = xs[i], ys[i], ts[i]
            T.add_trajectory(i, t)
        init = list(range(0, P, P // k))[:k]
        clusters = T.kmeansclustering(k, init=init, max_iter=iters)
        a_o, c_o, it_o = gp_ref.lloyd(xs, ys, ts, init, threshold=5.0, max_iter=iters)
        assert T.kmeans_iterations == it_o
        names = list(clusters)
        got = np.full(P, -1)
        for c, nm in enumerate(names):
            got[clusters[nm]] = c
        assert np.array_equal(got, a_o)
        cents = np.stack([np.stack([T.centroids[nm].xs for nm in names]), np.stack([T.centroids[nm].ys for nm in names]),
                          np.stack([T.centroids[nm].timestamp for nm in names])])
        assert np.array_equal(cents, c_o)


def test_config2_size_against_oracle_and_properties():
    # N=4096 fit compared with the oracle directly (about a second of CPU), prediction on a strided
    # sample of the 512x512 grid, plus size-independent properties of the full grid.
    X, Y, th = wl.single_path(4096, seed=2, D=2, R=2)
    m = GPmap.fit_gp(X, Y, theta=th)
    mo = gp_ref.fit(X, Y, th)
    assert nrm(np.tril(m.L.cpu().numpy()), mo["L"]) < 1e-9
    assert np.abs(m.lml - mo["lml"]).max() < LML_TOL * np.abs(mo["lml"]).max()
    Kd = dev(gp_ref.cov(X, th))
    assert float((Kd @ m.alpha - dev(Y)).abs().max()) < 1e-9            # K alpha = Y
    mu, var = m.predict_grid(wl.BOX, (512, 512))
    P = gp_ref.grid_points(wl.BOX, (512, 512))
    idx = np.arange(0, 512 * 512, 131)
    mu_o, var_o = gp_ref.predict(mo, P[idx])
    assert nrm(mu.view(-1, 2)[idx].cpu().numpy(), mu_o) < MEAN_TOL
    assert nrm(var.view(-1)[idx].cpu().numpy(), var_o) < VAR_TOL
    assert float(var.min()) > -1e-9 and float(var.max()) <= th[2] + 1e-12


def test_config4_size_factor_properties():
    # N=16384: L L^T reproduces K, and the factor agrees with cuSOLVER's (independent implementation)
    X, Y, th = wl.single_path(16384, seed=4, D=2, R=1)
    m = GPmap.fit_gp(X, Y, theta=th)
    Kfull = torch.from_numpy(gpu_cov(X, th)).cuda()
    L = m.L
    resid = float((L @ L.T - Kfull).abs().max())
    assert resid < 1e-12
    Lc = torch.linalg.cholesky(Kfull)
    assert float((L - Lc).abs().max()) / float(Lc.abs().max()) < 1e-9
    a_c = torch.cholesky_solve(dev(Y), Lc)
    assert float((m.alpha - a_c).abs().max()) / float(a_c.abs().max()) < 1e-6      # cond(K) ~ 1e6: alpha itself is loose
    lml_c = -0.5 * float((dev(Y) * a_c).sum()) - float(torch.log(torch.diagonal(Lc)).sum()) - 8192 * np.log(2 * np.pi)
    assert abs(m.lml[0] - lml_c) < LML_TOL * abs(lml_c)


def test_c_abi_argument_errors():
    lib = _native.load(); h = _native.handle(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    x = torch.zeros(8, 2, dtype=torch.float64, device="cuda"); k = torch.zeros(8, 8, dtype=torch.float64, device="cuda")
    th = _native.theta_array([1.0, 1.0, 1.0, 0.1])
    assert lib.gpm_cov(None, C.c_void_p(x.data_ptr()), 8, 2, th, C.c_void_p(k.data_ptr()), 8, 0, st) == -1
    assert lib.gpm_cov(h, C.c_void_p(x.data_ptr()), 8, 4, th, C.c_void_p(k.data_ptr()), 8, 0, st) == -5
    assert lib.gpm_cov(h, C.c_void_p(x.data_ptr()), 8, 2, th, C.c_void_p(k.data_ptr()), 7, 0, st) == -7
    assert b"argument 7" in lib.gpm_last_error()
    bad = _native.theta_array([1.0, -1.0, 1.0, 0.1])
    assert lib.gpm_cov(h, C.c_void_p(x.data_ptr()), 8, 2, bad, C.c_void_p(k.data_ptr()), 8, 0, st) == -5
    # gpm_fit: the one-call fit validates every argument before it launches anything
    p = lambda t: C.c_void_p(t.data_ptr())
    y = torch.zeros(8, 2, dtype=torch.float64, device="cuda"); a = torch.zeros_like(y)
    ws = torch.zeros(int(lib.gpm_potrf_workspace_bytes(8)) // 8, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda"); lml = torch.zeros(2, dtype=torch.float64, device="cuda")
    fit = lambda *v: lib.gpm_fit(*v)
    assert fit(None, p(x), 8, 2, th, p(y), 2, p(k), 8, p(ws), p(a), p(lml), p(info), st) == -1
    assert fit(h, p(x), 0, 2, th, p(y), 2, p(k), 8, p(ws), p(a), p(lml), p(info), st) == -3
    assert fit(h, p(x), 8, 2, bad, p(y), 2, p(k), 8, p(ws), p(a), p(lml), p(info), st) == -5
    assert fit(h, p(x), 8, 2, th, p(y), 9, p(k), 8, p(ws), p(a), p(lml), p(info), st) == -7
    assert fit(h, p(x), 8, 2, th, p(y), 2, p(k), 7, p(ws), p(a), p(lml), p(info), st) == -9
    assert fit(h, p(x), 8, 2, th, p(y), 2, p(k), 8, p(ws), p(y), p(lml), p(info), st) == -11      # alpha must not alias Y
    assert fit(h, p(x), 8, 2, th, p(y), 2, p(k), 8, p(ws), p(a), p(lml), None, st) == -13
    assert fit(h, p(x), 8, 2, th, p(y), 2, p(k), 8, p(ws), p(a), None, p(info), st) == 0           # lml is optional
    torch.cuda.synchronize()


def test_cross_cov_materialised_matches_oracle():
    lib = _native.load(); h = _native.handle(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for N, D, M in ((300, 2, 77), (130, 3, 513)):
        X, _, th = wl.single_path(N, seed=40 + N, D=D)
        rng = np.random.default_rng(N)
        Xs = np.column_stack([rng.uniform(-5e4, 5e4, M), rng.uniform(-5e4, 5e4, M)] + ([rng.uniform(0, 60, M)] if D == 3 else []))
        Xd, Xsd = dev(X), dev(Xs)
        ld = N + 3
        out = torch.full((M, ld), float("nan"), dtype=torch.float64, device="cuda")
        rc = lib.gpm_cross_cov(h, C.c_void_p(Xd.data_ptr()), N, D, _native.theta_array(th), C.c_void_p(Xsd.data_ptr()), None,
                               0, M, C.c_void_p(out.data_ptr()), ld, st)
        _native.check(rc, "gpm_cross_cov")
        want = gp_ref.cross_cov(X, Xs, th).T
        got = out.cpu().numpy()
        assert np.abs(got[:, :N] - want).max() < 4e-15
        assert np.isnan(got[:, N:]).all()                      # nothing written beyond column N
    # grid form: points [m0, m1) of a regular grid
    X, _, th = wl.single_path(200, seed=1, D=2)
    grid = _native.GpmGrid(wl.BOX[0], wl.BOX[1], wl.BOX[2], wl.BOX[3], 0.0, 31, 17)
    out = torch.empty((100, 200), dtype=torch.float64, device="cuda")
    Xd = dev(X)
    rc = lib.gpm_cross_cov(h, C.c_void_p(Xd.data_ptr()), 200, 2, _native.theta_array(th), None, C.byref(grid), 50, 150,
                           C.c_void_p(out.data_ptr()), 200, st)
    _native.check(rc, "gpm_cross_cov")
    P = gp_ref.grid_points(wl.BOX, (31, 17))[50:150]
    assert np.abs(out.cpu().numpy() - gp_ref.cross_cov(X, P, th).T).max() < 4e-15


def test_variance_chunked_workspace_and_per_step_path_agree():
    # a workspace far smaller than the query set forces the chunk loop; the per-step launch path
    # (GPM_VAR_STEPS) must give bitwise the same numbers as the fused persistent sweep
    lib = _native.load(); h = _native.handle(0)
    X, Y, th = wl.single_path(700, seed=77, D=2, R=2)
    m = GPmap.fit_gp(X, Y, theta=th)
    M = 5000
    rng = np.random.default_rng(5)
    Xs = dev(np.column_stack([rng.uniform(-5e4, 5e4, M), rng.uniform(-5e4, 5e4, M)]))
    mu_full, var_full = m.predict(Xs)

    def run(ws_rows):
        npad = 768
        ws = torch.empty(ws_rows * (npad + 1), dtype=torch.float64, device="cuda")
        mu = torch.empty((M, 2), dtype=torch.float64, device="cuda"); var = torch.empty(M, dtype=torch.float64, device="cuda")
        rc = lib.gpm_predict(h, C.c_void_p(m.X.data_ptr()), 700, 2, _native.theta_array(th), C.c_void_p(m.K.data_ptr()),
                             m.K.stride(0), C.c_void_p(m.ws.data_ptr()), C.c_void_p(m.alpha.data_ptr()), 2,
                             C.c_void_p(Xs.data_ptr()), None, 0, M, C.c_void_p(mu.data_ptr()), C.c_void_p(var.data_ptr()),
                             C.c_void_p(ws.data_ptr()), ws.numel() * 8, 3, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _native.check(rc, "gpm_predict")
        torch.cuda.synchronize()
        return mu, var

    mu_c, var_c = run(384)                      # 14 chunks of 384 rows
    assert torch.equal(mu_c, mu_full) and torch.equal(var_c, var_full)
    with _native.option("no_fused_mean", 1):    # separate mean kernel: same values up to summation order
        mu_n, var_n = run(384)
    assert torch.equal(var_n, var_full) and float((mu_n - mu_full).abs().max()) <= 1e-12 * float(mu_full.abs().max())
    with _native.option("var_steps", 1):
        mu_s, var_s = run(1024)
    assert torch.equal(var_s, var_full)
    mo = gp_ref.fit(X, Y, th)
    _, var_o = gp_ref.predict(mo, Xs.cpu().numpy())
    assert nrm(var_full.cpu().numpy(), var_o) < VAR_TOL
    # too small a workspace is an argument error, not a crash
    tiny = torch.empty(16, dtype=torch.float64, device="cuda")
    rc = lib.gpm_predict(h, C.c_void_p(m.X.data_ptr()), 700, 2, _native.theta_array(th), C.c_void_p(m.K.data_ptr()),
                         m.K.stride(0), C.c_void_p(m.ws.data_ptr()), C.c_void_p(m.alpha.data_ptr()), 2,
                         C.c_void_p(Xs.data_ptr()), None, 0, M, C.c_void_p(mu_c.data_ptr()), C.c_void_p(var_c.data_ptr()),
                         C.c_void_p(tiny.data_ptr()), tiny.numel() * 8, 3, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == -18


def test_batched_ragged_and_many_targets():
    rng = np.random.default_rng(9)
    for B, N, D, R in ((7, 33, 2, 1), (3, 257, 3, 8), (2, 640, 2, 3)):
        Xb = np.stack([wl.single_path(N, seed=500 + b, D=D)[0] for b in range(B)])
        Yb = rng.standard_normal((B, N, R))
        th = wl.default_theta(D)
        alpha, lml = GPmap.fit_gp_batched(Xb, Yb, theta=th)
        a_o, l_o = gp_ref.fit_batched(Xb, Yb, th)
        assert nrm(alpha.cpu().numpy(), a_o) < MEAN_TOL
        assert np.abs(lml.cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max()
    with pytest.raises(np.linalg.LinAlgError):
        GPmap.fit_gp_batched(np.zeros((2, 40, 2)), np.ones((2, 40, 1)), lengthscale=1.0, noise_var=0.0)


def test_export_raster_roundtrip(tmp_path):
    X, Y, th = wl.single_path(64, seed=3, D=2, R=2)
    m = GPmap.fit_gp(X, Y, theta=th)
    mu, var = m.predict_grid(wl.BOX, (12, 9))
    f = GPmap.export_raster(str(tmp_path / "r.npz"), mu, var, wl.BOX, (12, 9))
    z = np.load(f)
    assert z["mu"].shape == (9, 12, 2) and z["var"].shape == (9, 12) and z["x"].shape == (12,)
    assert np.array_equal(z["var"], var.cpu().numpy())


@pytest.mark.parametrize("N,D,R", [(33, 2, 1), (130, 3, 2), (300, 2, 8), (700, 3, 2)])
def test_lml_gradient_vs_oracle(N, D, R):
    X, Y2, th = wl.single_path(N, seed=60 + N, D=D, R=2)
    Y = np.random.default_rng(N).standard_normal((N, R))
    m = GPmap.fit_gp(X, Y, theta=th)
    g = m.lml_grad()
    go = gp_ref.lml_grad(X, Y, th)
    assert g.shape == (R, D + 2)
    assert np.abs(g - go).max() < 1e-7 * np.abs(go).max()


def test_lml_gradient_config2_size_finite_difference():
    # N=4096: compare the analytic gradient with central differences of the GPU LML itself
    X, Y, th = wl.single_path(4096, seed=2, D=2, R=2)
    g = GPmap.fit_gp(X, Y, theta=th).lml_grad()
    for j in (0, 2, 3):
        h = 1e-4
        tp, tm = th.copy(), th.copy()
        tp[j] *= np.exp(h); tm[j] *= np.exp(-h)
        fd = (GPmap.fit_gp(X, Y, theta=tp).lml - GPmap.fit_gp(X, Y, theta=tm).lml) / (2 * h)
        assert np.allclose(g[:, j], fd, rtol=1e-4, atol=1e-5 * np.abs(g).max())


def test_optimize_gp_improves_the_likelihood():
    X, Y, th = wl.single_path(400, seed=14, D=2, R=1)
    th0 = np.array([3000.0, 3000.0, 0.5, 0.5])
    m0 = GPmap.fit_gp(X, Y, theta=th0)
    theta, m, res = GPmap.optimize_gp(X, Y, th0, maxiter=25)
    assert m.lml.sum() > m0.lml.sum() + 1.0
    assert theta[0] == theta[1] and np.all(theta > 0)
    assert np.abs(m.lml_grad().sum(0)[2:]).max() < 1e-2 * max(1.0, abs(m.lml.sum()))


def test_cuda_graph_capture_replays_bitwise():
    # fit (look-ahead fork/join on the helper stream, cooperative solves) + predict captured once, replayed
    X, Y, th = wl.single_path(700, seed=21, D=2, R=2)
    Xd, Yd = dev(X), dev(Y)

    def step():
        m = GPmap.fit_gp(Xd, Yd, theta=th, check=False)
        mu, var = m.predict_grid(wl.BOX, (40, 30))
        return mu, var, m.alpha, m.lml_dev

    eager = [t.clone() for t in step()]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(3):
        Yd.mul_(1.0)                       # touch the input between replays
        graph.replay()
    torch.cuda.synchronize()
    for a, b in zip(out, eager):
        assert torch.equal(a, b)
    # new data through the same graph: static input buffers are re-read at replay
    Y2 = np.random.default_rng(3).standard_normal(Y.shape)
    Yd.copy_(dev(Y2))
    graph.replay(); torch.cuda.synchronize()
    mo = gp_ref.fit(X, Y2, th)
    assert nrm(out[2].cpu().numpy(), mo["alpha"]) < MEAN_TOL


def test_many_query_points_small_model():
    # 2.2 M query points in one chunk for a tiny model: more than 65535 row blocks in one launch
    X, Y, th = wl.single_path(64, seed=33, D=2, R=1)
    m = GPmap.fit_gp(X, Y, theta=th)
    mu, var = m.predict_grid(wl.BOX, (1500, 1480))
    mo = gp_ref.fit(X, Y, th)
    idx = np.arange(0, 1500 * 1480, 4099)
    P = gp_ref.grid_points(wl.BOX, (1500, 1480))[idx]
    mu_o, var_o = gp_ref.predict(mo, P)
    assert nrm(mu.view(-1, 1)[idx].cpu().numpy(), mu_o) < MEAN_TOL
    assert nrm(var.view(-1)[idx].cpu().numpy(), var_o) < VAR_TOL
    assert bool(torch.isfinite(var).all())


def test_model_checkpoint_roundtrip(tmp_path):
    X, Y, th = wl.single_path(300, seed=4, D=3, R=2)
    m = GPmap.fit_gp(X, Y, theta=th)
    mu, var = m.predict_grid(wl.BOX, (21, 17), t=5.0)
    f = tmp_path / "model.pt"
    torch.save(m.state_dict(), f)
    m2 = GPmap.GPModel.from_state_dict(torch.load(f, weights_only=False))
    mu2, var2 = m2.predict_grid(wl.BOX, (21, 17), t=5.0)
    assert torch.equal(mu, mu2) and torch.equal(var, var2) and np.array_equal(m.lml, m2.lml)


def test_batched_per_path_hyperparameters():
    Xb, Yb, th = wl.batched_paths(5, 200, seed=3, D=3, R=2)
    rng = np.random.default_rng(2)
    ths = np.stack([th * np.array([rng.uniform(0.5, 2), rng.uniform(0.5, 2), rng.uniform(0.5, 2), rng.uniform(0.5, 2), rng.uniform(0.3, 3)])
                    for _ in range(5)])
    alpha, lml = GPmap.fit_gp_batched(Xb, Yb, theta=ths)
    a_o, l_o = gp_ref.fit_batched(Xb, Yb, ths)
    assert nrm(alpha.cpu().numpy(), a_o) < MEAN_TOL
    assert np.abs(lml.cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max()


def test_run_to_run_determinism():
    # look-ahead streams, flag-chained solves and the multi-op sweep must not make results timing-dependent
    X, Y, th = wl.single_path(1500, seed=7, D=2, R=2)
    ref = None
    for _ in range(6):
        m = GPmap.fit_gp(X, Y, theta=th)
        mu, var = m.predict_grid(wl.BOX, (120, 97))
        cur = (m.alpha.clone(), m.lml_dev.clone(), mu.clone(), var.clone(), m.lml_grad())
        if ref is None:
            ref = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(ref[:4], cur[:4]))
            assert np.array_equal(ref[4], cur[4])


@pytest.mark.parametrize("path_fused", [1, 2], ids=["default-rule-tiled", "one-cta-per-path"])
def test_config3_size_sampled_against_oracle(path_fused):
    # 4096 paths x N=512 (BASELINE config 3), through both batched pipelines: every path factors (info = 0, finite
    # LML), a strided sample of paths matches the oracle, K alpha = Y holds for the sample, and the batch is invariant
    # to where a path sits in it (path b of the full batch == the same path fitted in a batch of 3).
    Xb, Yb, th = wl.batched_paths(4096, 512, seed=3, D=3, R=2)
    with _native.option("path_fused", path_fused):
        alpha, lml = GPmap.fit_gp_batched(Xb, Yb, theta=th)
    assert bool(torch.isfinite(lml).all()) and bool(torch.isfinite(alpha).all())
    idx = [0, 1, 777, 2048, 4095]
    a_o, l_o = gp_ref.fit_batched(Xb[idx], Yb[idx], th)
    assert nrm(alpha[idx].cpu().numpy(), a_o) < MEAN_TOL
    assert np.abs(lml[idx].cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max()
    for b in idx:
        K = gp_ref.cov(Xb[b], th)
        assert np.abs(K @ alpha[b].cpu().numpy() - Yb[b]).max() < 1e-9
    sub = [777, 4095, 0]
    a3, l3 = GPmap.fit_gp_batched(Xb[sub], Yb[sub], theta=th)
    # (small batches run the 16-warp diagonal-block kernel, large ones the 8-warp one: same maths, other summation order)
    assert nrm(a3.cpu().numpy(), alpha[sub].cpu().numpy()) < 1e-11
    assert np.abs((l3 - lml[sub]).cpu().numpy()).max() < 1e-12 * np.abs(l_o).max()


def test_config5_size_grid_slice_against_oracle():
    # N=16384 model (BASELINE config 5), one 2048-point slice of the 2048x2048 grid: the variance sweep at 128 block
    # columns against the oracle's triangular solve, and the sharded slice equals the same points of a wider slice.
    X, Y, th = wl.single_path(16384, seed=5, D=2, R=2)
    m = GPmap.fit_gp(X, Y, theta=th)
    mo = gp_ref.fit(X, Y, th)
    assert np.abs(m.lml - mo["lml"]).max() < LML_TOL * np.abs(mo["lml"]).max()
    lo = 1000 * 2048 + 512
    mu, var = m.predict_grid(wl.BOX, (2048, 2048), points=(lo, lo + 2048))
    P = gp_ref.grid_points(wl.BOX, (2048, 2048))[lo:lo + 2048]
    mu_o, var_o = gp_ref.predict(mo, P)
    assert nrm(mu.cpu().numpy(), mu_o) < MEAN_TOL
    assert nrm(var.cpu().numpy(), var_o) < VAR_TOL
    assert float(var.min()) > -1e-9 and float(var.max()) <= th[2] + 1e-12
    mu2, var2 = m.predict_grid(wl.BOX, (2048, 2048), points=(lo - 300, lo + 2048))
    assert torch.equal(mu2[300:], mu) and torch.allclose(var2[300:], var, rtol=0, atol=1e-13)


def test_separable_grid_kernels_match_the_pointwise_ones():
    """Grid queries use the separable form of the RBF kernel (PA + PB exponentials per patch and training point);
    the pointwise kernels (option no_separable, also what predict(Xs) runs) must agree to rounding, for ragged grids,
    D = 3 with a query time, R up to 8, and flat sub-ranges that start and end inside grid rows."""
    for (N, D, R, G, t) in ((300, 2, 2, (37, 23), None), (257, 3, 1, (5, 70), 31.0), (140, 2, 8, (130, 9), None)):
        X, Y, th = wl.single_path(N, seed=6, D=D, R=2)
        rng = np.random.default_rng(N)
        Y = np.ascontiguousarray(np.concatenate([Y, rng.standard_normal((N, 6))], axis=1)[:, :R])
        m = GPmap.fit_gp(X, Y, theta=th)
        M = G[0] * G[1]
        for pts in (None, (G[0] + 3, M - 5), (7, 9)):
            mu1, var1 = m.predict_grid(wl.BOX, G, t=t, points=pts)
            mo1 = m.predict_grid(wl.BOX, G, t=t, points=pts, return_var=False)
            with _native.option("no_separable", 1):
                mu0, var0 = m.predict_grid(wl.BOX, G, t=t, points=pts)
            assert torch.equal(mo1, mu1)                                   # mean-only and mean+variance calls agree bitwise
            assert nrm(mu1.cpu().numpy(), mu0.cpu().numpy()) < 1e-12
            assert nrm(var1.cpu().numpy(), var0.cpu().numpy()) < 1e-11
        # against the oracle, and against predict() on the same points given explicitly
        P = gp_ref.grid_points(wl.BOX, G, t=t)
        mo = gp_ref.fit(X, Y, th)
        mu_o, var_o = gp_ref.predict(mo, P)
        mu1, var1 = m.predict_grid(wl.BOX, G, t=t)
        assert nrm(mu1.reshape(-1, R).cpu().numpy(), mu_o) < MEAN_TOL
        assert nrm(var1.reshape(-1).cpu().numpy(), var_o) < VAR_TOL
        mu2, var2 = m.predict(P)
        assert nrm(mu1.reshape(-1, R).cpu().numpy(), mu2.cpu().numpy()) < 1e-12


def test_short_paths_one_cta_per_path_kernel():
    """N <= 112 (the reference resamples trajectories to 33 points, GPmap.py:189): the whole fit runs in one CTA per
    path.  Against the oracle, against the tiled pipeline (option no_small_fused), with per-path hyper-parameters, and
    the LAPACK-style info of a path that is not positive definite."""
    for (B, N, D, R) in ((7, 33, 2, 2), (3, 1, 2, 1), (4, 2, 3, 2), (5, 64, 3, 3), (3, 100, 2, 8), (6, 112, 3, 2)):
        Xb, Yb, th = wl.batched_paths(B, max(N, 4), seed=11, D=D, R=2)
        Xb, Yb = np.ascontiguousarray(Xb[:, :N]), Yb[:, :N]
        rng = np.random.default_rng(N)
        Yb = np.ascontiguousarray(np.concatenate([Yb, rng.standard_normal((B, N, 6))], axis=2)[:, :, :R])
        a1, l1 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
        with _native.option("no_small_fused", 1), _native.option("no_path_fused", 1):
            a0, l0 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
        a_o, l_o = gp_ref.fit_batched(Xb, Yb, th)
        assert nrm(a1.cpu().numpy(), a_o) < MEAN_TOL
        assert np.abs(l1.cpu().numpy() - l_o).max() < LML_TOL * max(np.abs(l_o).max(), 1.0)
        assert nrm(a1.cpu().numpy(), a0.cpu().numpy()) < 1e-10
        assert np.abs((l1 - l0).cpu().numpy()).max() < 1e-11 * max(np.abs(l_o).max(), 1.0)
    # per-path hyper-parameters
    Xb, Yb, th = wl.batched_paths(5, 33, seed=12, D=2, R=2)
    ths = np.stack([th * np.array([1.0 + 0.1 * b, 1.0 + 0.1 * b, 1.0 + 0.05 * b, 1.0 + b]) for b in range(5)])
    a1, l1 = GPmap.fit_gp_batched(Xb, Yb, theta=ths)
    a_o, l_o = gp_ref.fit_batched(Xb, Yb, ths)
    assert nrm(a1.cpu().numpy(), a_o) < MEAN_TOL
    assert np.abs(l1.cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max()
    # a duplicated sample with zero noise makes K singular: the path is reported, the others are unaffected
    Xs = Xb.copy(); Xs[2, 20] = Xs[2, 5]
    th0 = th.copy(); th0[-1] = 0.0
    with pytest.raises(np.linalg.LinAlgError, match="path 2"):
        GPmap.fit_gp_batched(Xs, Yb, theta=th0)
    # many more paths than gridDim.y allows for the tiled pipeline
    Xm, Ym, thm = wl.batched_paths(70000, 8, seed=13, D=2, R=1)
    am, lm = GPmap.fit_gp_batched(Xm, Ym, theta=thm)
    a_o, l_o = gp_ref.fit_batched(Xm[-3:], Ym[-3:], thm)
    assert nrm(am[-3:].cpu().numpy(), a_o) < MEAN_TOL and bool(torch.isfinite(lm).all())


def test_short_paths_second_row_threads():
    """A path of 33 .. 44 samples (the reference's 33) runs on ONE warp, the first N - 32 threads of which own a second
    row (fit_small_kernel<.., TWO>; option small_two_max).  Same arithmetic per element
    as one thread per row: alpha and info bitwise, the LML to rounding of its (re-associated) final sum; and against
    the oracle."""
    for (B, N, D, R) in ((9, 33, 2, 2), (5, 34, 3, 1), (4, 40, 2, 3), (3, 36, 3, 8), (3, 44, 2, 4), (2, 45, 2, 1), (2, 65, 3, 2)):
        Xb, Yb, th = wl.batched_paths(B, N, seed=21, D=D, R=2)
        rng = np.random.default_rng(N)
        Yb = np.ascontiguousarray(np.concatenate([Yb, rng.standard_normal((B, N, 6))], axis=2)[:, :, :R])
        a1, l1 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
        with _native.option("small_two_max", 0):
            a0, l0 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
        a_o, l_o = gp_ref.fit_batched(Xb, Yb, th)
        assert torch.equal(a1, a0)
        assert np.abs((l1 - l0).cpu().numpy()).max() < 1e-13 * max(np.abs(l_o).max(), 1.0)
        assert nrm(a1.cpu().numpy(), a_o) < MEAN_TOL
        assert np.abs(l1.cpu().numpy() - l_o).max() < LML_TOL * max(np.abs(l_o).max(), 1.0)
    # pivots that fail in second rows (rows 33..35 duplicate rows 5..7, zero noise: a tiny pivot of either sign, then a
    # negative one for certain), and in first rows
    Xb, Yb, th = wl.batched_paths(5, 40, seed=12, D=2, R=2)
    th0 = th.copy(); th0[-1] = 0.0
    for (row, path) in ((33, 1), (20, 3)):
        Xs = Xb.copy(); Xs[path, row:row + 3] = Xs[path, 5:8]
        with pytest.raises(np.linalg.LinAlgError, match=f"path {path}") as ei:
            GPmap.fit_gp_batched(Xs, Yb, theta=th0)
        msg1 = str(ei.value)
        with _native.option("small_two_max", 0):
            with pytest.raises(np.linalg.LinAlgError, match=f"path {path}") as ei0:
                GPmap.fit_gp_batched(Xs, Yb, theta=th0)
        assert msg1 == str(ei0.value)           # same pivot index


def test_latency_tile_kernel_is_bitwise_the_tile_kernel():
    """Launches of at most 37 tiles (every tile launch of a fit with N <= 4096, the tail of a large one) split each
    128x128 tile into four 64x64 quarters on four SMs (gemm_small.cu).  Same DMMA sequence per element: the factor must
    not change by a single bit, for full and ragged last blocks."""
    for N in (129, 900, 2048, 4096 + 77):
        X, _, th = wl.single_path(N, seed=7, D=2)
        Ko = gp_ref.cov(X, th)
        L1, ws1, info1 = gpu_potrf(Ko)
        with _native.option("no_small_tiles", 1):
            L0, ws0, info0 = gpu_potrf(Ko)
        assert info1 == 0 and info0 == 0
        assert np.array_equal(L1, L0)
        assert torch.equal(ws1, ws0)
        if N <= 2048:
            Lo = scipy.linalg.cholesky(Ko, lower=True)
            assert nrm(L1, Lo) < 1e-9


def test_two_streams_concurrent_fits_equal_the_serial_result_bitwise():
    """A library handle is single-stream state (helper stream, events, the flags of the chained solves); the Python
    binding keys its handles -- and its cached workspaces -- on (device, stream), so two host threads driving two
    CUDA streams at once cannot race.  Each must reproduce the serial result bit for bit."""
    import threading
    X, Y, th = wl.single_path(1500, seed=31, D=2, R=2)
    X2, Y2, th2 = wl.single_path(1100, seed=32, D=2, R=1)
    Xd, Yd, X2d, Y2d = dev(X), dev(Y), dev(X2), dev(Y2)
    shape = (96, 80)

    def work(Xa, Ya, tha):
        m = GPmap.fit_gp(Xa, Ya, theta=tha, check=False)
        mu, var = m.predict_grid(wl.BOX, shape)
        return m.alpha.clone(), m.lml_dev.clone(), mu.clone(), var.clone()

    want = [work(Xd, Yd, th), work(X2d, Y2d, th2)]
    torch.cuda.synchronize()
    n0 = len(_native._handles)
    got = [None, None]
    errs = []

    def runner(i, Xa, Ya, tha):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(4):
                    out = work(Xa, Ya, tha)
                s.synchronize()
            got[i] = out
        except Exception as e:                                   # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=runner, args=(0, Xd, Yd, th)), threading.Thread(target=runner, args=(1, X2d, Y2d, th2))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert len(_native._handles) >= n0 + 2                     # one handle per stream
    for w, g in zip(want, got):
        for a, b in zip(w, g):
            assert torch.equal(a, b)


def test_one_cta_per_path_tensor_core_fit():
    # option path_fused = 2 forces the one-CTA-per-path kernel whatever the batch size (the default rule gives batches
    # of more than one wave of CTAs to the tiled pipeline, which measures 2 - 6 % faster there)
    with _native.option("path_fused", 2):
        _one_cta_per_path_checks()
    # the default rule: a batch within one wave takes the one-CTA-per-path kernel, a larger one the tiled pipeline
    Xb, Yb, th = wl.batched_paths(600, 200, seed=43, D=2, R=1)
    a, l = GPmap.fit_gp_batched(Xb, Yb, theta=th)
    with _native.option("no_path_fused", 1):
        a0, l0 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
    assert torch.equal(a, a0) and torch.equal(l, l0)
    a, l = GPmap.fit_gp_batched(Xb[:5], Yb[:5], theta=th)
    with _native.option("path_fused", 2):
        a2, l2 = GPmap.fit_gp_batched(Xb[:5], Yb[:5], theta=th)
    assert torch.equal(a, a2) and torch.equal(l, l2)


def _one_cta_per_path_checks():
    """112 < N <= 1024 (BASELINE config 3): the whole fit of a path runs in one CTA -- covariance generated in
    registers, left-looking tile updates and solves on DMMA tiles, potf2 in shared memory, both substitutions and the
    LML (pathfit.cu).  Against the oracle and against the tiled batched pipeline (option no_path_fused), at ragged
    sizes around the 64-row tile and 128-column block edges, with R up to 8, more paths than resident CTAs (the
    per-CTA scratch is reused), per-path hyper-parameters and the LAPACK-style info of a singular path."""
    rng = np.random.default_rng(17)
    for (B, N, D, R) in ((3, 113, 2, 1), (2, 128, 3, 2), (3, 129, 2, 2), (2, 191, 3, 3), (2, 192, 2, 1), (3, 193, 2, 8),
                         (2, 320, 3, 2), (2, 384, 2, 4), (2, 1000, 2, 2), (1, 1024, 3, 1)):
        Xb = np.stack([wl.single_path(N, seed=900 + 7 * b + N, D=D)[0] for b in range(B)])
        Yb = rng.standard_normal((B, N, R))
        th = wl.default_theta(D)
        a1, l1 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
        with _native.option("no_path_fused", 1):
            a0, l0 = GPmap.fit_gp_batched(Xb, Yb, theta=th)
        a_o, l_o = gp_ref.fit_batched(Xb, Yb, th)
        assert nrm(a1.cpu().numpy(), a_o) < MEAN_TOL, (B, N, D, R)
        assert np.abs(l1.cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max(), (B, N, D, R)
        assert nrm(a1.cpu().numpy(), a0.cpu().numpy()) < 1e-10, (B, N, D, R)
        assert np.abs((l1 - l0).cpu().numpy()).max() < 1e-11 * np.abs(l_o).max(), (B, N, D, R)
    # more paths than resident CTAs (2 per SM): every path must equal the same path fitted alone, bit for bit
    Xb, Yb, th = wl.batched_paths(700, 200, seed=41, D=3, R=2)
    a, l = GPmap.fit_gp_batched(Xb, Yb, theta=th)
    for b in (0, 295, 296, 511, 699):
        a1, l1 = GPmap.fit_gp_batched(Xb[b:b + 1], Yb[b:b + 1], theta=th)
        assert torch.equal(a[b], a1[0]) and torch.equal(l[b], l1[0])
    a_o, l_o = gp_ref.fit_batched(Xb[-2:], Yb[-2:], th)
    assert nrm(a[-2:].cpu().numpy(), a_o) < MEAN_TOL
    # per-path hyper-parameters
    ths = np.stack([th * np.array([1.0 + 0.2 * b, 1.0 + 0.1 * b, 1.0 + 0.3 * b, 1.0 + 0.5 * b, 1.0 + b]) for b in range(4)])
    a, l = GPmap.fit_gp_batched(Xb[:4], Yb[:4], theta=ths)
    a_o, l_o = gp_ref.fit_batched(Xb[:4], Yb[:4], ths)
    assert nrm(a.cpu().numpy(), a_o) < MEAN_TOL and np.abs(l.cpu().numpy() - l_o).max() < LML_TOL * np.abs(l_o).max()
    # a duplicated sample with zero noise makes K singular: the path is reported, the others are unaffected
    # half of path 1's samples duplicated with zero noise: K has rank 100, some pivot must come out non-positive;
    # the path is reported (LAPACK-style info), the others are unaffected
    Xs = Xb[:3].copy(); Xs[1, 100:200] = Xs[1, 0:100]
    th0 = th.copy(); th0[-1] = 0.0
    with pytest.raises(np.linalg.LinAlgError, match="path 1"):
        GPmap.fit_gp_batched(Xs, Yb[:3], theta=th0)
    a2, _ = GPmap.fit_gp_batched(Xs, Yb[:3], theta=th0, check=False)
    a_ok, _ = GPmap.fit_gp_batched(Xs[[0, 2]], Yb[[0, 2]], theta=th0, check=False)
    assert torch.equal(a2[[0, 2]], a_ok)


@pytest.mark.parametrize("N,R", [(129, 1), (300, 2), (1000, 8), (2500, 2), (4173, 3), (6200, 2)])
def test_fit_with_fused_forward_substitution_equals_the_three_step_path(N, R):
    """gpm_fit (forward substitution riding on the factorisation: strip-kernel epilogues below ~4900 rows, the tile
    kernel's above) against cov + potrf + two chained solves (option no_fused_solve), and the solution itself
    against K alpha = Y.  The factor is bitwise the same; alpha differs by the summation order only."""
    X, Y, th = wl.single_path(N, seed=N, D=2, R=R)
    m1 = GPmap.fit_gp(X, Y, theta=th)
    with _native.option("no_fused_solve", 1):
        m0 = GPmap.fit_gp(X, Y, theta=th)
    assert torch.equal(m1.L, m0.L)
    assert nrm(m1.alpha.cpu().numpy(), m0.alpha.cpu().numpy()) < 1e-11
    assert np.abs(m1.lml - m0.lml).max() / np.abs(m0.lml).max() < 1e-12
    assert nrm(gp_ref.cov(X, th) @ m1.alpha.cpu().numpy(), Y.reshape(N, -1)) < 1e-9
    # deterministic: the same call again is bitwise the same
    m2 = GPmap.fit_gp(X, Y, theta=th)
    assert torch.equal(m1.alpha, m2.alpha) and torch.equal(m1.lml_dev, m2.lml_dev)


def test_fit_under_a_concurrent_sm_hogging_stream_is_bitwise_the_quiet_result():
    """The in-place panel solves (strip kernel: a CTA reads only the rows it writes) must not depend on how the block
    scheduler staggers their CTAs: the same fit while another stream keeps every SM busy with large GEMMs -- so that
    the fit's CTAs are admitted one by one as SMs free up -- must reproduce the quiet result bit for bit.  (Round 1's
    quarter-split kernel had an inter-CTA read-after-write race here that idle-GPU tests could not see.)"""
    X, Y, th = wl.single_path(2500, seed=77, D=2, R=2)
    m0 = GPmap.fit_gp(X, Y, theta=th)
    L0, a0, l0 = m0.L.clone(), m0.alpha.clone(), m0.lml_dev.clone()
    side = torch.cuda.Stream()
    A = torch.randn(8192, 8192, device="cuda")
    for trial in range(3):
        with torch.cuda.stream(side):
            for _ in range(4 + 2 * trial):
                B = A @ A                                    # noqa: F841  (tens of ms of SM-filling work)
        m1 = GPmap.fit_gp(X, Y, theta=th)
        torch.cuda.synchronize()
        assert torch.equal(m1.L, L0) and torch.equal(m1.alpha, a0) and torch.equal(m1.lml_dev, l0), trial
