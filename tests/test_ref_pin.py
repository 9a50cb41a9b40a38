"""Pins the oracle -- and the CUDA path -- to the REFERENCE's own code.

tests/golden/ref_vectors.npz holds input / output vectors produced by the reference's ObstacleTrack, InfiniteHorizonGP and
Matern32model sources, compiled from /root/reference against stand-in ROS / PCL / Eigen headers (oracle/_ref, see
oracle/ref_harness.cpp and tests/golden/make_ref_fixtures.py).  Three groups of tests:

  * test_golden_*      (CPU)  the oracle's restatements reproduce the reference's outputs on the committed vectors
  * test_live_*        (CPU)  fresh seeded inputs through oracle and oracle/_ref side by side; skipped where the compiled
                              reference is not present (it can only be built where /root/reference exists)
  * test_gpu_golden_*  (GPU)  the CUDA path through the C ABI reproduces the reference's outputs on the same vectors

Bit-exact for masks, ids, rings and circumcentres; rtol 1e-5 for the fp64 IHGP recursion (north_star's tolerance).
Not covered by these vectors: the PCL pieces (VoxelGrid, KdTree + EuclideanClusterExtraction, fromROSMsg), which are
third-party code outside the reference tree -- inside oracle/_ref they are the oracle's restatements.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "ref_vectors.npz"))


@pytest.fixture(scope="module")
def sim01():
    g = np.load(os.path.join(GOLD, "sim_01_occupancy.npz"))
    return g["occ"], np.float32(g["resolution"]), g["origin"]


def _trk_params(gold):
    return dict(zip([str(n) for n in gold["trk_param_names"]], gold["trk_params"].tolist()))


# ---------------------------------------------------------------------------------------------------------------------
# CPU: oracle vs the committed reference vectors
# ---------------------------------------------------------------------------------------------------------------------
def test_golden_ihgp_constants(oracle, gold):
    for dt, hyp, want in zip(gold["ihgp_dt"], gold["ihgp_hyp"], gold["ihgp_consts"]):
        got = oracle.ihgp_setup(float(dt), *hyp)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-13)  # closed-form expm / 2x2 inverse vs series / elimination


@pytest.mark.parametrize("tag", ["l10", "l40"])
def test_golden_ihgp_steps(oracle, gold, tag):
    rings, pv_ref, m_ref = gold[f"ihgp_{tag}_rings"], gold[f"ihgp_{tag}_pos_vel"], gold[f"ihgp_{tag}_m"]
    c = oracle.ihgp_setup(float(np.float32(0.1)), *gold["ihgp_hyp"][0])
    m = np.zeros(m_ref.shape[1:])
    for s in range(len(rings)):  # the GP means carry over from call to call, as inside the reference objects
        pv = oracle.ihgp_step(rings[s], m, 0.1, 0.03, c, c)
        np.testing.assert_allclose(pv, pv_ref[s], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m, m_ref[s], rtol=1e-9, atol=1e-12)
    assert (np.abs(pv_ref[..., 4:6]) == 1.5).any() and (np.abs(pv_ref[..., 4:6]) < 1.5).any()  # the clamp is exercised


def test_golden_remove_static(oracle, gold, sim01):
    occ, res, origin = sim01
    for quat, tol, pts, keep in zip(gold["rs_quat"], gold["rs_tol"], gold["rs_pts"], gold["rs_keep"]):
        kept, mask = oracle.remove_static(pts, occ, res, origin[:2], quat_xyzw=quat, static_tolerance=int(tol))
        assert np.array_equal(mask, keep)
        assert np.array_equal(kept, pts[keep.astype(bool)])


def test_golden_get_centroid(oracle, gold):
    stamp, time_init = gold["gc_stamp"]
    got = oracle.get_centroid(gold["gc_pts"], gold["gc_off"], gold["gc_idx"], stamp - time_init)
    assert np.array_equal(got.view(np.uint32), gold["gc_centroids"].view(np.uint32))


def _oracle_centroids(oracle, cloud, stamp, p, occ, res, origin):
    """clusterPointCloud restated (MOT.cpp:438-505): VoxelGrid -> removeStatic -> clustering -> getCentroid."""
    if len(cloud) == 0:
        return np.zeros((0, 4), np.float32)
    leaf = np.float32(p["voxel_leaf_size"])
    v = oracle.voxel_grid(cloud, (np.float32(1) * leaf, np.float32(1) * leaf, np.float32(20) * leaf))
    kept, _ = oracle.remove_static(v, occ, res, origin, static_tolerance=int(p["static_tolarance"]))
    if len(kept) == 0:
        return np.zeros((0, 4), np.float32)
    off, idx = oracle.cluster_kdtree(kept, p["cluster_tolerance"], int(p["min_cluster_size"]), int(p["max_cluster_size"]))
    return oracle.get_centroid(kept, off, idx, stamp)  # time_init = 0: stamps are below 1e9 (MOT.cpp:132-135)


def _trk_map(gold):
    h, w, res, ox, oy = gold["trk_map"]
    return np.zeros((int(h), int(w)), np.int8), np.float32(res), (ox, oy)


def test_golden_tracker_scenario(oracle, gold):
    from oracle.tracker_ref import TrackerRef
    p = _trk_params(gold)
    occ, res, origin = _trk_map(gold)
    hyp = tuple(gold["ihgp_hyp"][0])
    trk = TrackerRef(p["frequency"], p["id_threshold"], int(p["data_length"]), p["lpf_tau"], hyp, hyp)
    snap = {int(f): i for i, f in enumerate(gold["trk_track_frames"])}
    po, co, oo, to = gold["trk_pt_off"], gold["trk_cen_off"], gold["trk_out_off"], gold["trk_track_off"]
    for f, stamp in enumerate(gold["trk_stamps"]):
        cloud = gold["trk_pts"][po[f]:po[f + 1]]
        cen = _oracle_centroids(oracle, cloud, stamp, p, occ, res, origin)
        assert np.array_equal(cen.view(np.uint32), gold["trk_centroids"][co[f]:co[f + 1]].view(np.uint32))
        r = trk.step(cen, stamp)
        assert (r is not None) == bool(gold["trk_produced"][f])
        if r is not None:
            assert np.array_equal(r[0], gold["trk_ids"][oo[f]:oo[f + 1]])
            want = gold["trk_pos_vel"][oo[f]:oo[f + 1]]
            assert np.array_equal(r[1][:, :2], want[:, :2])                            # LPF position: float arithmetic, exact
            np.testing.assert_allclose(r[1][:, 4:6], want[:, 4:6], rtol=1e-6, atol=1e-7)  # IHGP velocity
        if f in snap:
            i = snap[f]
            sl = slice(to[i], to[i + 1])
            assert np.array_equal(np.array(trk.obj_ids, np.int32), gold["trk_track_ids"][sl])
            if trk.obj_ids:
                assert np.array_equal(np.array(trk.stack, np.float32).view(np.uint32), gold["trk_track_rings"][sl].view(np.uint32))
                np.testing.assert_allclose(np.array(trk.m), gold["trk_track_m"][sl], rtol=1e-9, atol=1e-12)
    assert trk.next_obj_num > 16 and len(trk.obj_ids) < trk.next_obj_num  # births after frame 0 and the 5 s purge both happened


def test_golden_obstacle_rows(oracle, gold):
    # publishObstacles (MOT.cpp:253-295): radius 0.3, covariance[0] = .1 on every obstacle of every published message
    pv = gold["trk_pos_vel"]
    assert np.all(pv[:, 3] == np.float32(0.3)) and np.all(pv[:, 7] == np.float32(0.1)) and np.all(pv[:, 2] == 0) and np.all(pv[:, 6] == 0)
    rows = oracle.obstacle_table(np.c_[pv[:, :2], np.zeros((len(pv), 2), np.float32), pv[:, 4:6], np.zeros((len(pv), 2), np.float32)], gold["trk_ids"])
    assert np.array_equal(rows[:, 0], gold["trk_ids"]) and np.all(rows[:, 1] == np.float32(0.3)) and np.all(rows[:, 6] == np.float32(0.1))


# ---------------------------------------------------------------------------------------------------------------------
# CPU: oracle vs the compiled reference, live (development container only)
# ---------------------------------------------------------------------------------------------------------------------
def _live():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libmot_ref.so is not built here (needs /root/reference)")
    return ref


@pytest.mark.parametrize("seed", range(4))
def test_live_ihgp(oracle, seed):
    ref = _live()
    rng = np.random.default_rng(seed)
    hyp = (float(np.exp(rng.uniform(-7, -3))), float(np.exp(rng.uniform(-5, 0))), float(np.exp(rng.uniform(-0.5, 1.5))))
    freq = float(rng.choice([5.0, 10.0, 20.0]))
    L, T = int(rng.integers(4, 30)), 16
    dt = float(np.float32(1) / np.float32(freq))
    c_ref = ref.ihgp_constants(dt, *hyp)
    c = oracle.ihgp_setup(dt, *hyp)
    np.testing.assert_allclose(c, c_ref, rtol=1e-9, atol=1e-13)
    R = ref.Reference(frequency=freq, lpf_tau=0.05, data_length=L, logSigma2_x=np.log(hyp[0]), logMagnSigma2_x=np.log(hyp[1]),
                      logLengthScale_x=np.log(hyp[2]), logSigma2_y=np.log(hyp[0]), logMagnSigma2_y=np.log(hyp[1]), logLengthScale_y=np.log(hyp[2]))
    # the reference exponentiates the log parameters again (MOT.cpp:524-530): feed the oracle the same round trip
    hyp_rt = tuple(float(np.exp(np.log(h))) for h in hyp)
    c = oracle.ihgp_setup(dt, *hyp_rt)
    m = np.zeros((T, 4))
    for s in range(4):
        rings = np.zeros((T, L, 4), np.float32)
        rings[:, :, :2] = rng.uniform(-5, 5, (T, 1, 2)) + np.cumsum(rng.normal(0, 0.05, (T, L, 2)), axis=1)
        rings[:, :, 3] = 50.0 + dt * (np.arange(L) + s)
        pv_ref, m_ref = R.call_ihgp(rings)
        pv = oracle.ihgp_step(rings, m, np.float32(dt), 0.05, c, c)
        np.testing.assert_allclose(pv, pv_ref, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m, m_ref, rtol=1e-9, atol=1e-12)
    R.close()


@pytest.mark.parametrize("seed", range(4))
def test_live_remove_static_and_yaw(oracle, synth, seed):
    ref = _live()
    rng = np.random.default_rng(100 + seed)
    occ, res, origin = synth.make_map_c1(cells=400)
    occ = occ.copy()
    occ[rng.integers(40, 360, 300), rng.integers(40, 360, 300)] = 100   # scattered occupied cells
    occ[rng.integers(40, 360, 100), rng.integers(40, 360, 100)] = -1    # and unknown ones
    occ[rng.integers(40, 360, 100), rng.integers(40, 360, 100)] = 50    # 50 is NOT above the threshold (> 50)
    occ[rng.integers(40, 360, 100), rng.integers(40, 360, 100)] = 51
    yaw = float(rng.uniform(-np.pi, np.pi)) if seed else 0.0
    tol = int(rng.integers(0, 5))
    quat = np.array([0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2)])
    H, W = occ.shape
    n = 5000
    u = rng.uniform((tol + 1) * res, (W - tol - 1) * res, n)
    v = rng.uniform((tol + 1) * res, (H - tol - 1) * res, n)
    pts = np.ones((n, 4), np.float32)
    pts[:, 0] = origin[0] + np.cos(yaw) * u - np.sin(yaw) * v
    pts[:, 1] = origin[1] + np.sin(yaw) * u + np.cos(yaw) * v
    pts[:, 2] = rng.uniform(0, 2, n)
    R = ref.Reference(static_tolarance=tol)
    R.set_map(occ, res, origin[:2], quat)
    assert np.float32(R.yaw_from_quat(quat)) == np.float32(oracle.yaw_from_quat(quat))
    kept_ref = R.remove_static(pts)
    kept, _ = oracle.remove_static(pts, occ, res, origin[:2], quat_xyzw=quat, static_tolerance=tol)
    assert np.array_equal(kept, kept_ref) and 0 < len(kept) < n
    # a point whose window leaves the map: undefined behaviour in the reference (the stand-in reports it), dropped by the oracle
    out = np.array([[origin[0] - 1.0, origin[1] - 1.0, 0.5, 1.0]], np.float32)
    R.set_map(occ, res, origin[:2])
    with pytest.raises(IndexError):
        R.remove_static(out)
    assert len(oracle.remove_static(out, occ, res, origin[:2], static_tolerance=tol)[0]) == 0
    R.close()


@pytest.mark.parametrize("seed", range(3))
def test_live_get_centroid(oracle, synth, seed):
    ref = _live()
    cloud, _ = synth.make_frame_c1(n_points=16384, frame=seed)
    occ, res, origin = synth.make_map_c1()
    kept, _ = oracle.remove_static(cloud, occ, res, origin[:2])
    off, idx = oracle.cluster_kdtree(kept, 0.3, 3, 300)
    R = ref.Reference()
    want = R.get_centroid(kept, off, idx, stamp=77.0 + seed, time_init=70.0)
    got = oracle.get_centroid(kept, off, idx, 7.0 + seed)
    assert len(off) - 1 >= 10 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    R.close()


def test_live_cloud_callback(oracle, gold):
    # the whole node callback, a different seed than the committed scenario
    import importlib.util
    ref = _live()
    from oracle.tracker_ref import TrackerRef
    spec = importlib.util.spec_from_file_location("make_ref_fixtures", os.path.join(GOLD, "make_ref_fixtures.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    p = dict(gen.TRK_PARAMS)
    occ, res, origin = _trk_map(gold)
    hyp = tuple(gold["ihgp_hyp"][0])
    R = ref.Reference(**p)
    R.set_map(occ, res, origin)
    trk = TrackerRef(p["frequency"], p["id_threshold"], p["data_length"], p["lpf_tau"], hyp, hyp)
    produced = 0
    for stamp, cloud in gen.tracker_scenario(5, n_frames=100, n_obj=10):
        r_ref = R.cloud_callback(cloud, stamp)
        r = trk.step(_oracle_centroids(oracle, cloud, stamp, p, occ, res, origin), stamp)
        assert (r is None) == (r_ref is None)
        if r is not None:
            produced += 1
            assert np.array_equal(r[0], r_ref[0]) and np.array_equal(r[1][:, :2], r_ref[1][:, :2])
            np.testing.assert_allclose(r[1][:, 4:6], r_ref[1][:, 4:6], rtol=1e-6, atol=1e-7)
        ids, rings, m = R.tracks()
        assert np.array_equal(ids, np.array(trk.obj_ids, np.int32))
        if len(ids):
            assert np.array_equal(rings.view(np.uint32), np.array(trk.stack, np.float32).view(np.uint32))
            np.testing.assert_allclose(m, np.array(trk.m), rtol=1e-9, atol=1e-12)
    assert produced > 60
    R.close()


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path (through the C ABI) vs the committed reference vectors
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_golden_ihgp(mot, gold):
    t = mot.Tracker(device=0, max_points=1024, max_tracks=64)
    for dt, hyp, want in zip(gold["ihgp_dt"], gold["ihgp_hyp"], gold["ihgp_consts"]):
        t.ihgp_configure(float(dt), 0.03, tuple(hyp), tuple(hyp), 10)
        np.testing.assert_allclose(t.ihgp_constants(0), want, rtol=1e-9, atol=1e-13)
    for tag, L in (("l10", 10), ("l40", 40)):
        rings, pv_ref, m_ref = gold[f"ihgp_{tag}_rings"], gold[f"ihgp_{tag}_pos_vel"], gold[f"ihgp_{tag}_m"]
        hyp = tuple(gold["ihgp_hyp"][0])
        t.ihgp_configure(0.1, 0.03, hyp, hyp, L)
        m = np.zeros(m_ref.shape[1:])
        for s in range(len(rings)):
            pv = t.ihgp_step(rings[s], m)
            np.testing.assert_allclose(pv, pv_ref[s], rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(m, m_ref[s], rtol=RTOL, atol=1e-9)
    t.close()


@pytest.mark.gpu
def test_gpu_golden_remove_static(mot, gold, sim01):
    occ, res, origin = sim01
    t = mot.Tracker(device=0, max_points=1 << 14, max_tracks=0)
    for quat, tol, pts, keep in zip(gold["rs_quat"], gold["rs_tol"], gold["rs_pts"], gold["rs_keep"]):
        t.set_map(occ, res, origin[:2], quat_xyzw=tuple(quat), static_tolarance=int(tol))
        assert np.array_equal(t.remove_static(pts), pts[keep.astype(bool)])
    t.close()


@pytest.mark.gpu
def test_gpu_golden_tracker_scenario(mot, gold):
    # the whole SURVEY 8 path on the device -- VoxelGrid, removeStatic, clustering, circumcentres, association, IHGP,
    # obstacle rows -- against what the reference's cloudCallback produced for the same frames
    p = _trk_params(gold)
    occ, res, origin = _trk_map(gold)
    L = int(p["data_length"])
    hyp = tuple(gold["ihgp_hyp"][0])
    t = mot.Tracker(device=0, max_points=1 << 12, max_tracks=512)
    t.set_map(occ, res, origin, static_tolarance=int(p["static_tolarance"]))
    t.set_cluster_params(p["cluster_tolerance"], int(p["min_cluster_size"]), int(p["max_cluster_size"]))
    t.ihgp_configure(float(np.float32(1) / np.float32(p["frequency"])), p["lpf_tau"], hyp, hyp, L)
    leaf = np.float32(p["voxel_leaf_size"])
    snap = {int(f): i for i, f in enumerate(gold["trk_track_frames"])}
    po, co, oo, to = gold["trk_pt_off"], gold["trk_cen_off"], gold["trk_out_off"], gold["trk_track_off"]
    for f, stamp in enumerate(gold["trk_stamps"]):
        cloud = gold["trk_pts"][po[f]:po[f + 1]]
        cen = np.zeros((0, 4), np.float32)
        if len(cloud):
            kept = t.remove_static(t.voxel_grid(cloud, (leaf, leaf, np.float32(20) * leaf)))
            if len(kept):
                t.extract(kept)
                cen = t.get_centroid(stamp)
        assert np.array_equal(cen.view(np.uint32), gold["trk_centroids"][co[f]:co[f + 1]].view(np.uint32))
        out = t.tracks_step(cen, stamp, p["id_threshold"], p["frequency"])
        assert out["produced"] == bool(gold["trk_produced"][f])
        if out["produced"]:
            want = gold["trk_pos_vel"][oo[f]:oo[f + 1]]
            assert np.array_equal(out["ids"], gold["trk_ids"][oo[f]:oo[f + 1]])
            np.testing.assert_allclose(out["pos_vel"][:, :2], want[:, :2], rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(out["pos_vel"][:, 4:6], want[:, 4:6], rtol=RTOL, atol=1e-6)
            ob = out["obstacles"]
            assert np.array_equal(ob["id"], gold["trk_ids"][oo[f]:oo[f + 1]]) and np.all(ob["radius"] == want[:, 3])
        if f in snap:
            i = snap[f]
            sl = slice(to[i], to[i + 1])
            ids, rings, m = t.tracks_get()
            assert np.array_equal(ids, gold["trk_track_ids"][sl])
            if len(ids):
                assert np.array_equal(rings.view(np.uint32), gold["trk_track_rings"][sl].view(np.uint32))
                np.testing.assert_allclose(m, gold["trk_track_m"][sl], rtol=RTOL, atol=1e-9)
    t.close()
