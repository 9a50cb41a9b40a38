"""GPU tests of the single-launch small-frame path (csrc/frame_small.cuh, mot_small_frames): the whole per-frame path of a frame of
the size the reference tracker sees, in one kernel.  Every result is compared with the oracle (bit-exact kept cloud, partition and
CSR; rtol 1e-5 statistics and circumcentres) and with the general path on the same handle; the hand-back conditions are provoked
one by one and must give the same results through the general path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _frame_vs_oracle(t, oracle, cloud, occ, resn, origin, p, stamp=2.5):
    out = t.frame(cloud, stamp)
    kept_ref, _ = oracle.remove_static(cloud, occ, resn, origin[:2], static_tolerance=p["static_tolerance"])
    off_ref, idx_ref = oracle.cluster_kdtree(kept_ref, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    assert np.array_equal(out["kept"], kept_ref), "kept cloud"
    assert np.array_equal(out["offsets"], off_ref) and np.array_equal(out["indices"], idx_ref), "CSR"
    cen_ref = oracle.get_centroid(kept_ref, off_ref, idx_ref, stamp)
    np.testing.assert_allclose(out["centroids"], cen_ref, rtol=RTOL, atol=1e-6)
    for c in range(len(off_ref) - 1):
        pts = kept_ref[idx_ref[off_ref[c]:off_ref[c + 1]], :3]
        st = out["stats"][c]
        assert st["count"] == len(pts)
        np.testing.assert_allclose(st["mean"], pts.mean(0, dtype=np.float64), rtol=RTOL, atol=1e-5)
        assert np.array_equal(st["bbox_min"], pts.min(0)) and np.array_equal(st["bbox_max"], pts.max(0))
    lab = oracle.labels_grid(kept_ref, p["cluster_tolerance"]) if len(kept_ref) else np.zeros(0, np.int32)
    assert np.array_equal(t.result_labels(), lab), "component labels"
    return out


def test_c1_frame_takes_the_small_path(mot, oracle, synth):
    occ, resn, origin = synth.make_map_c1()
    p = synth.C1_PARAMS
    t = mot.Tracker(device=0, max_points=1 << 17, max_tracks=0)
    ctas, h0, m0 = t.small_frames()
    assert ctas in (8, 16) and (h0, m0) == (0, 0)
    t.set_map(occ, resn, origin[:2], static_tolarance=p["static_tolerance"])
    t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    for f in range(4):
        cloud, _ = synth.make_frame_c1(frame=f, box_shift=(0.4 * f, -0.3 * f))
        out = _frame_vs_oracle(t, oracle, cloud, occ, resn, origin, p, stamp=0.1 * f)
        assert out["K"] > 0
        assert t.last_launches() == 5    # front, cell pairs, tables, farthest pair, finish -- one graph, one host round trip
    assert t.small_frames()[1:] == (4, 0)
    # the same frames through the general path on the same handle: identical arrays
    cloud, _ = synth.make_frame_c1(frame=2)
    a = t.frame(cloud, 1.0)
    t.small_frames(0)
    b = t.frame(cloud, 1.0)
    assert t.last_launches() > 10
    for k in ("kept", "offsets", "indices"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["centroids"], b["centroids"])   # the same arithmetic, bit for bit
    for k in ("count", "bbox_min", "bbox_max"):
        assert np.array_equal(a["stats"][k], b["stats"][k])
    np.testing.assert_allclose(a["stats"]["mean"], b["stats"]["mean"], rtol=1e-6)
    t.close()


def test_extract_then_late_centroids(mot, oracle, synth):
    # ec.extract() through the small path, getCentroid afterwards (the adapter's order of calls, MOT.cpp:488-491)
    p = synth.C1_PARAMS
    cloud, _ = synth.make_frame_c1(n_points=30000)
    t = mot.Tracker(device=0, max_points=1 << 16, max_tracks=0)
    t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    off, idx = t.extract(cloud)
    assert t.small_frames()[1:] == (1, 0)
    o_ref, i_ref = oracle.cluster_kdtree(cloud, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    assert np.array_equal(off, o_ref) and np.array_equal(idx, i_ref)
    cen = t.get_centroid(3.0)
    np.testing.assert_allclose(cen, oracle.get_centroid(cloud, o_ref, i_ref, 3.0), rtol=RTOL, atol=1e-6)
    st = t.cluster_stats()
    assert np.array_equal(st["count"], np.diff(o_ref))
    t.close()


@pytest.mark.parametrize("n", [1, 2, 31, 33, 1000, 1025, 16385, 70001, 131072])
def test_sizes_and_shapes(mot, oracle, n):
    rng = np.random.default_rng(n)
    # a few blobs + background noise, ragged sizes around the CTA / cluster boundaries
    k = max(1, n // 400)
    centres = rng.uniform(-30, 30, (k, 3)).astype(np.float32)
    pts = centres[rng.integers(0, k, n)] + rng.normal(0, 0.25, (n, 3)).astype(np.float32)
    noise = rng.random(n) < 0.1
    pts[noise] = rng.uniform(-40, 40, (int(noise.sum()), 3)).astype(np.float32)
    cloud = np.zeros((n, 4), np.float32)
    cloud[:, :3] = pts
    t = mot.Tracker(device=0, max_points=max(n, 64), max_tracks=0)
    for tol, mn, mx in ((0.2, 1, 1 << 30), (0.35, 3, 500)):
        t.set_cluster_params(tol, mn, mx)
        off, idx = t.extract(cloud)
        lab = oracle.labels_grid(cloud, tol)
        o_ref, i_ref = oracle.csr_from_labels(lab, mn, mx)
        assert np.array_equal(t.result_labels(), lab)
        assert np.array_equal(off, o_ref) and np.array_equal(idx, i_ref)
    t.close()


def test_hand_back_conditions(mot, oracle, synth):
    t = mot.Tracker(device=0, max_points=1 << 16, max_tracks=0)
    rng = np.random.default_rng(5)

    def run(cloud, tol, mn, mx):
        t.small_frames(131072)           # also clears the back-off that follows a hand-back
        before = t.small_frames()
        t.set_cluster_params(tol, mn, mx)
        off, idx = t.extract(cloud)
        lab = oracle.labels_grid(cloud, tol)
        o_ref, i_ref = oracle.csr_from_labels(lab, mn, mx)
        assert np.array_equal(t.result_labels(), lab)
        assert np.array_equal(off, o_ref) and np.array_equal(idx, i_ref)
        after = t.small_frames()
        return after[1] - before[1], after[2] - before[2]

    def cloud_of(xyz):
        c = np.zeros((len(xyz), 4), np.float32)
        c[:, :3] = xyz
        return c

    # 1. crowded cells: 6000 points inside eight 0.25 m fine cells -> above the per-cell bound of the witness search
    dense = cloud_of(rng.uniform(0.0, 0.45, (6000, 3)))
    assert run(dense, 0.5, 1, 1 << 30) == (0, 1)
    # 2. more clusters than the counting rank takes: 5000 isolated points, min size 1
    g = np.stack(np.meshgrid(np.arange(50), np.arange(50), np.arange(2), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 2.0
    assert run(cloud_of(g), 0.5, 1, 1 << 30) == (0, 1)
    # ... the same points with min size 2: no cluster at all, served by the small path
    assert run(cloud_of(g), 0.5, 2, 1 << 30) == (1, 0)
    # 3. one cluster too large for the O(size^2) ordering: a chain of 6000 points, 0.1 apart
    chain = np.zeros((6000, 3), np.float32)
    chain[:, 0] = np.arange(6000) * 0.1
    assert run(cloud_of(chain), 0.15, 1, 1 << 30) == (0, 1)
    # ... dropped by max size: served
    assert run(cloud_of(chain), 0.15, 1, 100) == (1, 0)
    # 4. coordinates beyond the cell key (|x| / tol > 2^20)
    far = cloud_of(rng.uniform(-1, 1, (500, 3)))
    far[7, 0] = 4.0e5
    assert run(far, 0.3, 1, 1 << 30) == (0, 1)
    # after a hand-back the next calls go straight to the general path (back-off), then the small path is tried again
    t.small_frames(131072)
    ok = cloud_of(rng.uniform(-5, 5, (2000, 3)))
    t.set_cluster_params(0.5, 1, 1 << 30)
    t.extract(dense)
    h0, m0 = t.small_frames()[1:]
    for _ in range(15):
        t.extract(ok)
    assert t.small_frames()[1:] == (h0, m0)
    t.extract(ok)
    assert t.small_frames()[1:] == (h0 + 1, m0)
    # NaN is the documented error on either path, and the handle keeps working
    bad = ok.copy()
    bad[11, 1] = np.nan
    with pytest.raises(mot.MotError):
        t.extract(bad)
    off, idx = t.extract(ok)
    o_ref, i_ref = oracle.csr_from_labels(oracle.labels_grid(ok, 0.5), 1, 1 << 30)
    assert np.array_equal(off, o_ref) and np.array_equal(idx, i_ref)
    t.close()


def test_everything_removed_and_empty(mot, oracle, synth):
    occ, resn, origin = synth.make_map_c1()
    p = synth.C1_PARAMS
    cloud, _ = synth.make_frame_c1(n_points=5000)
    t = mot.Tracker(device=0, max_points=1 << 14, max_tracks=0)
    t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    t.set_map(np.full_like(occ, 100), resn, origin[:2], static_tolarance=0)
    out = t.frame(cloud, 0.0)
    assert out["m"] == 0 and out["K"] == 0 and len(out["indices"]) == 0 and np.array_equal(out["offsets"], [0])
    assert t.small_frames()[1:] == (1, 0)
    out = t.frame(np.zeros((0, 4), np.float32), 0.0)
    assert out["m"] == 0 and out["K"] == 0
    t.close()


def test_pointcloud2_through_the_small_path(mot, oracle, synth):
    occ, resn, origin = synth.make_map_c1()
    p = synth.C1_PARAMS
    cloud, _ = synth.make_frame_c1(n_points=40000)
    raw = np.zeros((len(cloud), 8), np.float32)     # point_step 32: x y z pad intensity ring pad pad
    raw[:, :3] = cloud[:, :3]
    t = mot.Tracker(device=0, max_points=1 << 16, max_tracks=0)
    t.set_map(occ, resn, origin[:2], static_tolarance=p["static_tolerance"])
    t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    out = t.cluster_pointcloud2(raw.view(np.uint8).reshape(-1), len(cloud), 32, (0, 4, 8), do_remove_static=True, stamp_minus_time_init=1.5)
    assert t.small_frames()[1:] == (1, 0)
    kept_ref, _ = oracle.remove_static(cloud, occ, resn, origin[:2], static_tolerance=p["static_tolerance"])
    o_ref, i_ref = oracle.cluster_kdtree(kept_ref, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    assert np.array_equal(out["offsets"], o_ref) and np.array_equal(out["indices"], i_ref)
    np.testing.assert_allclose(out["centroids"], oracle.get_centroid(kept_ref, o_ref, i_ref, 1.5), rtol=RTOL, atol=1e-6)
    t.close()


@pytest.mark.parametrize("seed", range(12))
def test_chains_and_random_frames(mot, oracle, seed):
    # shapes that stress the shared-memory union-find of k_fs_tables: long thin components (a spiral and a zig-zag whose points sit
    # just under the tolerance apart: thousands of cells in ONE component, joined through chains of cell pairs), planes, blobs and
    # noise, at sizes up to the path's limit of kept points; the small path must take every frame and agree with the oracle
    rng = np.random.default_rng(1000 + seed)
    tol = float(rng.choice([0.12, 0.2, 0.35, 0.5, 0.8]))
    n_chain = int(rng.integers(500, 12000))
    s = np.arange(n_chain, dtype=np.float64) * (tol * rng.uniform(0.55, 0.98))
    ang = s / 6.0
    spiral = np.stack([(4 + 0.3 * ang) * np.cos(ang), (4 + 0.3 * ang) * np.sin(ang), 0.02 * s], 1)
    zz = np.stack([s % 37.0, 50 + (s // 37.0) * tol * 0.9, np.zeros_like(s)], 1)
    n_plane = int(rng.integers(0, 9000))
    plane = np.stack([rng.uniform(-60, -40, n_plane), rng.uniform(-10, 10, n_plane), np.full(n_plane, 1.5)], 1)
    k = int(rng.integers(1, 40))
    centres = rng.uniform(-30, 30, (k, 3))
    n_blob = int(rng.integers(100, 15000))
    blobs = centres[rng.integers(0, k, n_blob)] + rng.normal(0, tol * 0.8, (n_blob, 3))
    n_noise = int(rng.integers(0, 3000))
    noise = rng.uniform(-80, 80, (n_noise, 3))
    pts = np.concatenate([spiral, zz, plane, blobs, noise]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    assert len(pts) <= 49152
    cloud = np.zeros((len(pts), 4), np.float32)
    cloud[:, :3] = pts
    t = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
    t.small_frames(max_points=131072)
    mn, mx = int(rng.choice([1, 3, 10])), int(rng.choice([200, 5000, 1 << 30]))
    t.set_cluster_params(tol, mn, mx)
    before = t.small_frames()
    off, idx = t.extract(cloud)
    after = t.small_frames()
    lab = oracle.labels_grid(cloud, tol)
    o_ref, i_ref = oracle.csr_from_labels(lab, mn, mx)
    assert np.array_equal(t.result_labels(), lab), "component labels"
    assert np.array_equal(off, o_ref) and np.array_equal(idx, i_ref), "CSR"
    took_small = after[1] == before[1] + 1
    crowded = np.unique(np.floor(pts.astype(np.float64) * (2.0 / (tol * (1 + 2.0 ** -10)))).astype(np.int64), axis=0, return_counts=True)[1].max() > 64
    many = len(o_ref) - 1 > 4096
    assert took_small or crowded or many or (after[2] > before[2]), "the frame neither took the small path nor was handed back"
    t.close()
