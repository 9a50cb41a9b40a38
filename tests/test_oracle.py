"""CPU tests of the oracle itself: the restated reference path (KD-tree + BFS) against two independent
implementations of the same fp32 predicate, hand-built known answers, scipy cross-checks and the one
reference-owned fixture (map/sim_01).  The reference has no tests or golden vectors; the pieces of
the oracle that restate the reference's own code are pinned in tests/test_ref_pin.py, the PCL pieces tested here stay
"unpinned" with respect to real PCL."""
import os

import numpy as np
import pytest

from cases import kat_cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(kat_cases().keys()))
def test_kat_three_way(oracle, name):
    pts, tol, mn, mx, expect_k = kat_cases()[name]
    lab_b = oracle.labels_bruteforce(pts, tol)
    lab_g = oracle.labels_grid(pts, tol)
    assert np.array_equal(lab_b, lab_g)
    off_b, idx_b = oracle.csr_from_labels(lab_b, mn, mx)
    off_k, idx_k = oracle.cluster_kdtree(pts, tol, mn, mx)
    assert np.array_equal(off_b, off_k) and np.array_equal(idx_b, idx_k)
    if expect_k is not None:
        assert len(off_k) - 1 == expect_k
    sizes = np.diff(off_k)
    assert np.all(sizes[:-1] >= sizes[1:])
    for c in range(len(sizes)):
        seg = idx_k[off_k[c]:off_k[c + 1]]
        assert np.all(seg[1:] > seg[:-1])


def test_scipy_cross_check(oracle):
    # float64 KD-tree + connected components; points are snapped to a 1/64 lattice and tol sits between lattice
    # distances, so fp32-vs-fp64 rounding cannot flip any pair (guard band)
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(0)
    xyz = np.round(rng.uniform(0, 6, (5000, 3)) * 64) / 64
    pts = np.ones((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    tol = 0.2
    pairs = cKDTree(xyz).query_pairs(tol, output_type="ndarray")
    d = np.linalg.norm(xyz[pairs[:, 0]] - xyz[pairs[:, 1]], axis=1)
    assert np.all(np.abs(d - tol) > 1e-4)
    n = len(xyz)
    g = coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    ncomp, comp = connected_components(g, directed=False)
    lab = oracle.labels_bruteforce(pts, tol)
    assert len(np.unique(lab)) == ncomp
    first = np.full(ncomp, n)
    np.minimum.at(first, comp, np.arange(n))
    assert np.array_equal(lab, first[comp])


def test_random_clouds_kdtree_vs_bruteforce(oracle, synth):
    occ, res, origin = synth.make_map_c1()
    cloud, _ = synth.make_frame_c1(n_points=32768)
    kept, keep = oracle.remove_static(cloud, occ, res, origin[:2])
    assert 3000 < len(kept) < 20000
    lab = oracle.labels_bruteforce(kept, 0.3)
    off_b, idx_b = oracle.csr_from_labels(lab, 5, 300)
    off_k, idx_k = oracle.cluster_kdtree(kept, 0.3, 5, 300)
    assert np.array_equal(off_b, off_k) and np.array_equal(idx_b, idx_k)
    assert len(off_k) - 1 >= 10


def test_remove_static_sim01_fixture(oracle):
    g = np.load(os.path.join(GOLD, "sim_01_occupancy.npz"))
    occ, res, origin = g["occ"], float(g["resolution"]), g["origin"]
    assert occ.shape == (214, 93)
    vals, counts = np.unique(occ, return_counts=True)
    assert dict(zip(vals.tolist(), counts.tolist())) == {-1: 3301, 0: 15407, 100: 1194}
    # independent numpy restatement: dilate the blocked mask, look the cell up
    t = 2
    blocked = (occ > 50) | (occ == -1)
    pad = np.pad(blocked, t, constant_values=True)
    dil = np.zeros_like(blocked)
    for i in range(2 * t + 1):
        for j in range(2 * t + 1):
            dil |= pad[i:i + occ.shape[0], j:j + occ.shape[1]]
    assert int((~dil).sum()) == 14223  # passable cells after the 5x5 dilation
    rng = np.random.default_rng(1)
    n = 20000
    pts = np.ones((n, 4), np.float32)
    pts[:, 0] = rng.uniform(origin[0] - 0.3, origin[0] + 93 * res + 0.3, n)
    pts[:, 1] = rng.uniform(origin[1] - 0.3, origin[1] + 214 * res + 0.3, n)
    kept, keep = oracle.remove_static(pts, occ, res, origin[:2], static_tolerance=t)
    xm = (pts[:, 0].astype(np.float64) - origin[0]).astype(np.float32)
    ym = (pts[:, 1].astype(np.float64) - origin[1]).astype(np.float32)
    col = np.trunc(xm / np.float32(res)).astype(np.int64)
    row = np.trunc(ym / np.float32(res)).astype(np.int64)
    inside = (col >= 0) & (col < 93) & (row >= 0) & (row < 214)
    ref = np.zeros(n, bool)
    ref[inside] = ~dil[row[inside], col[inside]]
    assert np.array_equal(keep.astype(bool), ref)
    assert 0.2 < ref.mean() < 0.9


def test_ihgp_constants_table(oracle):
    # SURVEY 8a-4 table, launch hyper-parameters (computed there with numpy/scipy)
    c = oracle.ihgp_setup(0.1, np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
    np.testing.assert_allclose(c[0:4], [0.99683012, 0.09214412, -0.0616804, 0.84605232], atol=5e-9)
    np.testing.assert_allclose(c[8:10], [0.29829458, 0.52246828], atol=5e-9)
    np.testing.assert_allclose(c[4:8], [0.6994811, 0.06465803, -0.58249252, 0.79790994], atol=5e-9)
    np.testing.assert_allclose(c[10:14], [0.94655783, -0.07640631, 0.97361388, 0.55083324], atol=5e-9)
    np.testing.assert_allclose(c[14], 0.0058240557, atol=5e-11)


def test_ihgp_against_scipy(oracle):
    from scipy.linalg import expm, solve_discrete_are

    sigma2, magn, ell, dt = np.exp(-5.0), np.exp(-3.0), np.exp(0.5), 0.1
    lam = np.sqrt(3) / ell
    F = np.array([[0, 1], [-lam ** 2, -2 * lam]])
    Pinf = np.diag([magn, magn * lam ** 2])
    H = np.array([[1.0, 0.0]])
    A = expm(F * dt)
    Q = Pinf - A @ Pinf @ A.T
    PP = solve_discrete_are(A.T, H.T, Q, np.array([[sigma2]]))
    S = (H @ PP @ H.T)[0, 0] + sigma2
    K = (PP @ H.T / S)[:, 0]
    PF = PP - np.outer(K, H @ PP)
    AKHA = A - np.outer(K, H @ A)
    G = np.linalg.solve(A @ PF @ A.T + Q, A @ PF).T
    c = oracle.ihgp_setup(dt, sigma2, magn, ell)
    np.testing.assert_allclose(c[0:4], A.ravel(), rtol=1e-12)
    np.testing.assert_allclose(c[4:8], AKHA.ravel(), rtol=1e-7)   # fixed-point DARE stops at 1e-10 (IHGP.cpp:9)
    np.testing.assert_allclose(c[8:10], K, rtol=1e-7)
    np.testing.assert_allclose(c[10:14], G.ravel(), rtol=1e-7)
    # step response: numpy restatement of update()/getEft() with these matrices
    L = 12
    rings = np.zeros((1, L, 4), np.float32)
    rings[0, :, 0] = np.linspace(0, 1.1, L) ** 2
    rings[0, :, 1] = -0.5 * np.arange(L) * 0.1
    rings[0, :, 3] = np.arange(L) * 0.1
    m = np.zeros((1, 4))
    pv = oracle.ihgp_step(rings, m, dt, 0.03, c, c)
    for axis in range(2):
        v = ((rings[0, 1:, axis] - rings[0, :-1, axis]) / np.float32(dt)).astype(np.float64)
        mean = v.mean()
        mm = np.zeros(2)
        MF = []
        for y in v - mean:
            mm = AKHA @ mm + K * y
            MF.append(mm)
        out = MF[-1][0] + mean
        ms = MF[-1]
        for k in range(len(MF) - 2, -1, -1):
            ms = MF[k] + G @ (ms - A @ MF[k])
        np.testing.assert_allclose(pv[0, 4 + axis], np.clip(out, -1.5, 1.5), rtol=1e-5)
        np.testing.assert_allclose(m[0, 2 * axis:2 * axis + 2], ms, rtol=1e-5, atol=1e-9)


def test_get_centroid_known_answers(oracle):
    # right triangle: circumcentre is the midpoint of the hypotenuse
    pts = np.ones((3, 4), np.float32)
    pts[:, :3] = [[0, 0, 0], [4, 0, 0], [0, 3, 0]]
    off, idx = np.array([0, 3], np.int32), np.array([0, 1, 2], np.int32)
    c = oracle.get_centroid(pts, off, idx, 7.5)
    np.testing.assert_allclose(c[0], [2.0, 1.5, 0.0, 7.5], atol=1e-6)
    # vertical farthest pair: slope is +-inf, every line distance is NaN -> Pk stays zero (UB policy)
    pts[:, :3] = [[1, 0, 0], [1, 5, 0], [2, 2, 0]]
    c = oracle.get_centroid(pts, off, idx, 0.0)
    assert np.all(np.isfinite(c[0]))
    # two coincident points: G == 0 -> centroid = Pi
    pts2 = np.ones((2, 4), np.float32)
    pts2[:, :3] = [[3, 4, 1], [3, 4, 1]]
    c = oracle.get_centroid(pts2, np.array([0, 2], np.int32), np.array([0, 1], np.int32), 0.0)
    np.testing.assert_allclose(c[0, :2], [3, 4])


def test_voxel_grid(oracle):
    rng = np.random.default_rng(2)
    xyz = rng.uniform(-3, 3, (5000, 3)).astype(np.float32)
    pts = np.ones((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    leaf = (0.1, 0.1, 2.0)
    out = oracle.voxel_grid(pts, leaf)
    ijk = np.floor(xyz / np.array(leaf, np.float32)).astype(np.int64)
    ijk -= ijk.min(0)
    dims = ijk.max(0) + 1
    lin = ijk[:, 0] + ijk[:, 1] * dims[0] + ijk[:, 2] * dims[0] * dims[1]
    u, inv = np.unique(lin, return_inverse=True)
    assert len(out) == len(u)
    ref = np.zeros((len(u), 3))
    np.add.at(ref, inv, xyz.astype(np.float64))
    ref /= np.bincount(inv)[:, None]
    np.testing.assert_allclose(out[:, :3], ref, rtol=1e-5, atol=1e-5)


def test_tracker_reference_known_answers():
    # hand-checkable behaviour of the restated association (MOT.cpp:176-219, 593-619, 545-584)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.tracker_ref import TrackerRef
    hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
    trk = TrackerRef(10.0, 0.4, 5, 0.03, hyp, hyp)

    def cen(rows):
        return np.array(rows, dtype=np.float32).reshape(-1, 4)

    assert trk.step(cen([]), 0.0) is None and trk.first                      # empty first frame: nothing happens
    assert trk.step(cen([[0, 0, 0, 1.0], [5, 0, 0, 1.0]]), 1.0) is None        # first frame registers, no output
    assert trk.obj_ids == [0, 1] and len(trk.stack[0]) == 5
    ids, pv = trk.step(cen([[5.1, 0, 0, 1.1], [0.1, 0, 0, 1.1], [9, 9, 0, 1.1]]), 1.1)
    assert ids.tolist() == [1, 0, 2]                                            # first match in registration order; new id for the stranger
    assert trk.stack[1][-1][0] == np.float32(5.1) and trk.stack[1][0][0] == np.float32(5.0)
    # two centroids inside id_threshold of track 0 -> both take id 0 (non exclusive), track filtered twice
    ids, _ = trk.step(cen([[0.15, 0, 0, 1.2], [0.2, 0.05, 0, 1.2]]), 1.2)
    assert ids.tolist() == [0, 0]
    # a 0.5 s gap (> 3 dt) on track 1 -> 4 interpolated centroids + the new one: the ring (L = 5) is fully rewritten
    ids, _ = trk.step(cen([[5.35, 0, 0, 1.6]]), 1.6)
    assert ids.tolist() == [1]
    ring = np.array(trk.stack[1])
    np.testing.assert_allclose(ring[:, 3], [1.2, 1.3, 1.4, 1.5, 1.6], atol=1e-5)
    np.testing.assert_allclose(np.diff(ring[:4, 0]), (5.35 - 5.1) / 4, atol=1e-5)
    # nothing is purged before 5 * frequency callbacks; afterwards tracks unseen for > 5 s go
    for k in range(60):
        trk.step(cen([[0.2, 0.0, 0, 2.0 + 0.1 * k]]), 2.0 + 0.1 * k)
    assert trk.obj_ids == [0]
