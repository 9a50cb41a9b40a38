"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same inputs.

Bars: bit-exact for the removeStatic output, the cluster partition and the CSR arrays (integer / index work);
rtol 1e-5 for centroids, bbox/mean and IHGP states (north_star's fp32 tolerance)."""
import os

import numpy as np
import pytest

from cases import kat_cases

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module", params=["cell-path", "auto", "small"])
def trk(mot, request):
    # every test that takes `trk` runs three times: with the cell-based union-find of round 2 forced (MOT_UF_MODE=2; by default
    # it only serves calls of >= 1.5 M points), with the default choice of the general path (the round-1 sweep kernels at these
    # sizes) -- both with the single-launch small-frame path switched off --, and with the library's defaults, where single
    # frames of up to 131,072 points take the small-frame path (k_frame_small) and fall back to the general one when they must
    saved = {k: os.environ.get(k) for k in ("MOT_UF_MODE", "MOT_SMALL_POINTS")}
    if request.param == "cell-path":
        os.environ["MOT_UF_MODE"] = "2"
    if request.param != "small":
        os.environ["MOT_SMALL_POINTS"] = "0"
    try:
        t = mot.Tracker(device=0, max_points=1 << 21, max_tracks=2048)   # the switches are read by mot_create
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    t.variant = request.param
    yield t
    t.close()


def check_extract(trk, oracle, pts, tol, mn, mx, brute=True):
    trk.set_cluster_params(tol, mn, mx)
    off, idx = trk.extract(pts)
    if brute:
        lab = oracle.labels_bruteforce(pts, tol)
    else:
        lab = oracle.labels_grid(pts, tol)
    off_ref, idx_ref = oracle.csr_from_labels(lab, max(mn, 1), mx)
    assert np.array_equal(trk.result_labels(), lab), "component labels differ"
    assert np.array_equal(off, off_ref), "cluster offsets differ"
    assert np.array_equal(idx, idx_ref), "point indices differ"
    return off, idx


@pytest.mark.parametrize("name", sorted(kat_cases().keys()))
def test_kat(trk, oracle, name):
    pts, tol, mn, mx, expect_k = kat_cases()[name]
    off, idx = check_extract(trk, oracle, pts, tol, mn, mx)
    if expect_k is not None:
        assert len(off) - 1 == expect_k
    # the reference path itself (KD-tree + BFS restatement) agrees as well
    off_kd, idx_kd = oracle.cluster_kdtree(pts, tol, max(mn, 1), mx)
    assert np.array_equal(off, off_kd) and np.array_equal(idx, idx_kd)


def test_remove_static_sim01(trk, oracle):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sim_01_occupancy.npz"))
    occ, res, origin = g["occ"], float(g["resolution"]), g["origin"]
    rng = np.random.default_rng(7)
    H, W = occ.shape
    n = 50000
    xy = np.c_[rng.uniform(origin[0] - 0.5, origin[0] + W * res + 0.5, n), rng.uniform(origin[1] - 0.5, origin[1] + H * res + 0.5, n)]
    pts = np.ones((n, 4), np.float32)
    pts[:, :2] = xy
    pts[:, 2] = rng.uniform(0, 2, n)
    for t in (0, 2, 4):
        for quat in ((0, 0, 0, 1), (0, 0, np.sin(0.35), np.cos(0.35))):
            trk.set_map(occ, res, origin[:2], quat_xyzw=quat, static_tolarance=t)
            kept = trk.remove_static(pts)
            ref, keep = oracle.remove_static(pts, occ, res, origin[:2], quat_xyzw=quat, static_tolerance=t)
            assert 0 < len(ref) < n
            assert np.array_equal(kept, ref)


def test_remove_static_large_map_and_edges(trk, oracle, synth):
    # a bitmap too large for shared memory (global lookups) and points outside / on the border of the map
    occ = np.zeros((1200, 1100), np.int8)
    occ[::37, :] = 100
    occ[:, ::53] = -1
    rng = np.random.default_rng(11)
    pts = np.ones((30000, 4), np.float32)
    pts[:, 0] = rng.uniform(-5, 60, len(pts))
    pts[:, 1] = rng.uniform(-5, 65, len(pts))
    pts[:100, 0] = np.nan
    pts[100:200, 1] = 1e30
    trk.set_map(occ, 0.05, (0.0, 0.0), static_tolarance=1)
    kept = trk.remove_static(pts)
    ref, _ = oracle.remove_static(pts, occ, 0.05, (0.0, 0.0), static_tolerance=1)
    assert np.array_equal(kept, ref)
    assert len(trk.remove_static(np.zeros((0, 4), np.float32))) == 0


def test_c1_frame_fused(trk, oracle, synth):
    occ, res, origin = synth.make_map_c1()
    cloud, _ = synth.make_frame_c1()
    p = synth.C1_PARAMS
    trk.set_map(occ, res, origin[:2], static_tolarance=p["static_tolerance"])
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    out = trk.frame(cloud, stamp_minus_time_init=12.25)
    kept_ref, _ = oracle.remove_static(cloud, occ, res, origin[:2], static_tolerance=p["static_tolerance"])
    assert np.array_equal(out["kept"], kept_ref)
    off_ref, idx_ref = oracle.cluster_kdtree(kept_ref, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    assert np.array_equal(out["offsets"], off_ref) and np.array_equal(out["indices"], idx_ref)
    assert out["K"] >= 10
    cen_ref = oracle.get_centroid(kept_ref, off_ref, idx_ref, 12.25)
    np.testing.assert_allclose(out["centroids"], cen_ref, rtol=RTOL, atol=1e-6)
    st_ref = oracle.cluster_stats(kept_ref, off_ref, idx_ref)
    st = out["stats"]
    assert np.array_equal(st["count"], st_ref[:, 0].astype(np.int32))
    np.testing.assert_allclose(st["mean"], st_ref[:, 1:4], rtol=RTOL, atol=1e-6)
    assert np.array_equal(st["bbox_min"], st_ref[:, 4:7]) and np.array_equal(st["bbox_max"], st_ref[:, 7:10])


def test_centroid_bitexact_small_clusters(trk, oracle):
    # getCentroid is deterministic IEEE arithmetic in both implementations: expect identical bits
    rng = np.random.default_rng(3)
    parts = [np.asarray(c) + rng.normal(0, 0.15, size=(rng.integers(5, 120), 3)) for c in rng.uniform(-30, 30, size=(60, 3))]
    pts = np.ones((sum(len(p) for p in parts), 4), np.float32)
    pts[:, :3] = np.concatenate(parts)
    trk.set_cluster_params(0.35, 3, 5000)
    off, idx = trk.extract(pts)
    cen = trk.get_centroid(3.0)
    ref = oracle.get_centroid(pts, off, idx, 3.0)
    assert len(off) - 1 > 20
    assert np.array_equal(cen.view(np.uint32), ref.view(np.uint32))


def test_c2_slice_vs_kdtree(trk, oracle, synth):
    # a 64k azimuth wedge of the c2 frame, against the reference path (KD-tree + BFS)
    p = synth.C2_PARAMS
    frame = synth.scene_c2().frame(0)
    wedge = frame[np.arctan2(frame[:, 1], frame[:, 0]) < -np.pi + 2 * np.pi / 16]
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    off, idx = trk.extract(wedge)
    off_ref, idx_ref = oracle.cluster_kdtree(wedge, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    assert np.array_equal(off, off_ref) and np.array_equal(idx, idx_ref)


def test_c2_full_frame(trk, oracle, synth):
    p = synth.C2_PARAMS
    frame = synth.scene_c2().frame(0)
    assert len(frame) == 1 << 20
    off, idx = check_extract(trk, oracle, frame, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"], brute=False)
    # size-independent properties: indices ascending inside clusters, sizes descending, a partition of a subset
    sizes = np.diff(off)
    assert np.all(sizes[:-1] >= sizes[1:]) and sizes.min() >= p["min_cluster_size"] and sizes.max() <= p["max_cluster_size"]
    assert len(np.unique(idx)) == len(idx)
    for c in range(0, len(off) - 1, 17):
        seg = idx[off[c]:off[c + 1]]
        assert np.all(seg[1:] > seg[:-1])
    st = trk.cluster_stats()
    st_ref = oracle.cluster_stats(frame, off, idx)
    np.testing.assert_allclose(st["mean"], st_ref[:, 1:4], rtol=RTOL, atol=1e-5)
    assert np.array_equal(st["bbox_min"], st_ref[:, 4:7]) and np.array_equal(st["bbox_max"], st_ref[:, 7:10])
    # idempotence: clustering the same frame again gives the same bytes
    off2, idx2 = trk.extract(frame)
    assert np.array_equal(off, off2) and np.array_equal(idx, idx2)


@pytest.mark.parametrize("tol", [0.1, 0.3, 1.0])
def test_c4_blobs(trk, oracle, synth, tol):
    p = synth.C4_PARAMS
    frame = synth.make_frame_c4(n_points=1 << 18, n_blobs=125)
    check_extract(trk, oracle, frame, tol, p["min_cluster_size"], p["max_cluster_size"], brute=False)


def test_batch_matches_single_frames(trk, oracle, synth):
    p = synth.C3_PARAMS
    sc = synth.scene_c3()
    frames = [sc.frame(f, n_points=20000 + 1000 * f) for f in range(5)] + [np.zeros((0, 4), np.float32)]
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    fco, off, idx = trk.extract_batch(frames)
    assert fco[0] == 0 and fco[-1] == len(off) - 1
    st_batch = trk.cluster_stats()
    assert len(st_batch) == len(off) - 1
    for f, fr in enumerate(frames):
        if len(fr):
            k0, k1 = fco[f], fco[f + 1]
            ref = oracle.cluster_stats(fr, off[k0:k1 + 1] - off[k0], idx[off[k0]:off[k1]])
            assert np.array_equal(st_batch["count"][k0:k1], ref[:, 0].astype(np.int32))
            np.testing.assert_allclose(st_batch["mean"][k0:k1], ref[:, 1:4], rtol=RTOL, atol=1e-6)
            assert np.array_equal(st_batch["bbox_min"][k0:k1], ref[:, 4:7]) and np.array_equal(st_batch["bbox_max"][k0:k1], ref[:, 7:10])
    for f, fr in enumerate(frames):
        o1, i1 = trk.extract(fr)
        k0, k1 = fco[f], fco[f + 1]
        assert k1 - k0 == len(o1) - 1
        assert np.array_equal(off[k0:k1 + 1] - off[k0], o1)
        assert np.array_equal(idx[off[k0]:off[k1]], i1)


def test_concurrent_handles_on_one_gpu(mot, oracle, synth):
    # several handles (one host thread + streams each) share the GPU, as in bench.py and in a multi-LiDAR deployment:
    # every handle must produce exactly what it produces alone
    import threading
    p = synth.C3_PARAMS
    sc = synth.scene_c3()
    n_handles, rounds = 4, 6
    frames = [sc.frame(10 + f, n_points=60000 + 5000 * f) for f in range(n_handles)]
    refs = []
    for fr in frames:
        lab = oracle.labels_grid(fr, p["cluster_tolerance"])
        refs.append(oracle.csr_from_labels(lab, p["min_cluster_size"], p["max_cluster_size"]))
    handles = [mot.Tracker(device=0, max_points=1 << 17, max_tracks=0) for _ in range(n_handles)]
    errors = []

    def worker(i):
        try:
            t = handles[i]
            t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
            for r in range(rounds):
                j = (i + r) % n_handles   # every handle sees every frame, in a different order
                off, idx = t.extract(frames[j])
                if not (np.array_equal(off, refs[j][0]) and np.array_equal(idx, refs[j][1])):
                    errors.append((i, r, j))
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    th = [threading.Thread(target=worker, args=(i,)) for i in range(n_handles)]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join()
    for t in handles:
        t.close()
    assert not errors, errors


ALTERNATIVES = [{"MOT_UF_MODE": "2"}, {"MOT_UF_MODE": "2", "MOT_UF_XMODE": "1"}, {"MOT_UF_MODE": "2", "MOT_UF_XMODE": "0"},
                {"MOT_UF_MODE": "2", "MOT_UF_XMODE": "1", "MOT_UF_PHASES": "7,24"}, {"MOT_UF_MODE": "2", "MOT_UF_FBLOCKS": "2"}, {"MOT_UF_MODE": "1"}, {"MOT_UF_MODE": "1", "MOT_UF_PAIR": "0"}, {"MOT_UF_MODE": "0"}, {"MOT_UF_MODE": "2", "MOT_UF_LIGHT": "1"}, {"MOT_UF_MODE": "2", "MOT_UF_LIGHT": "1024"},
                {"MOT_UF_MODE": "2", "MOT_UF_XMODE": "0", "MOT_UF_SPLIT": "1"}, {"MOT_UF_MODE": "1", "MOT_UF_BLOCKS": "3"}, {"MOT_UF_MODE": "1", "MOT_UF_TMA": "0"}, {"MOT_SORT_MODE": "1"}, {"MOT_SORT_BITS": "8"}, {"MOT_KEYS_HIST": "1"}, {"MOT_CSR_COMPACT": "0"}, {"MOT_CSR_COMPACT": "8"}, {"MOT_SORT_BIGTILE": "1000"},
                {"MOT_UF_PRIO": "0"}, {"MOT_SYNC": "yield"}, {"MOT_SYNC": "block"}, {"MOT_PLAN_SPEC": "0"},
                {"MOT_SMALL_POINTS": "131072"}, {"MOT_SMALL_POINTS": "131072", "MOT_SMALL_CLUSTER": "8"}, {"MOT_SMALL_POINTS": "131072", "MOT_SMALL_CLUSTER": "2"},
                {"MOT_SMALL_POINTS": "131072", "MOT_SMALL_CELLCAP": "3"}, {"MOT_HOST_STAGE": "0"}, {"MOT_SMALL_POINTS": "131072", "MOT_SMALL_GRAPH": "0"}]


@pytest.mark.parametrize("env", ALTERNATIVES, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_alternative_paths_keep_parity(mot, oracle, synth, env):
    # every A/B switch (README) selects a different kernel or host path for the same result: known answers, one LiDAR
    # frame slice with dense and sparse neighbourhoods, one frame batch
    env = dict(env)
    env.setdefault("MOT_SMALL_POINTS", "0")   # the switches of the general path are what is under test, unless the entry says otherwise
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        t = mot.Tracker(device=0, max_points=1 << 19, max_tracks=0)   # the switches are read by mot_create
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for name in ("bridged_blobs", "pair_exactly_tol", "pair_tol_minus_ulp", "duplicates", "uniform_dense", "size_filter_edges", "empty"):
        if name in kat_cases():
            pts, tol, mn, mx, _ = kat_cases()[name]
            check_extract(t, oracle, pts, tol, mn, mx)
    p = synth.C2_PARAMS
    frame = synth.scene_c2().frame(3)
    az = np.arctan2(frame[:, 1], frame[:, 0])
    wedge = np.ascontiguousarray(frame[np.abs(az) < 0.6])
    check_extract(t, oracle, wedge, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"], brute=False)
    p3 = synth.C3_PARAMS
    sc3 = synth.scene_c3()
    frames = [sc3.frame(40 + f, n_points=30000) for f in range(3)]
    t.set_cluster_params(p3["cluster_tolerance"], p3["min_cluster_size"], p3["max_cluster_size"])
    fco, off, idx = t.extract_batch(frames)
    for f, fr in enumerate(frames):
        lab = oracle.labels_grid(fr, p3["cluster_tolerance"])
        o_ref, i_ref = oracle.csr_from_labels(lab, p3["min_cluster_size"], p3["max_cluster_size"])
        k0, k1 = fco[f], fco[f + 1]
        assert np.array_equal(off[k0:k1 + 1] - off[k0], o_ref) and np.array_equal(idx[off[k0]:off[k1]], i_ref)
    t.close()


def test_u64_key_path_is_taken(trk, oracle):
    pts, tol, mn, mx, _ = kat_cases()["wide_extent_u64_keys"]
    check_extract(trk, oracle, pts, tol, mn, mx)
    if trk.variant != "small":   # (the small-frame path hashes absolute cell coordinates: no sort key at all)
        assert trk.result_grid()["key_bits"] > 32
    # and in batch mode (frame id in the top key bits)
    trk.set_cluster_params(tol, mn, mx)
    fco, off, idx = trk.extract_batch([pts, pts[::-1].copy(), pts[:500]])
    o1, i1 = trk.extract(pts[::-1].copy())
    assert np.array_equal(off[fco[1]:fco[2] + 1] - off[fco[1]], o1) and np.array_equal(idx[off[fco[1]]:off[fco[2]]], i1)


def test_many_clusters_large_k_path(trk, oracle):
    # > 8192 clusters forces the 64-bit radix ordering path
    rng = np.random.default_rng(5)
    centres = np.stack(np.meshgrid(np.arange(110), np.arange(100), indexing="ij"), -1).reshape(-1, 2) * 2.0
    pts = np.ones((len(centres) * 3, 4), np.float32)
    pts[:, :2] = np.repeat(centres, 3, axis=0) + rng.uniform(-0.1, 0.1, (len(centres) * 3, 2))
    pts[:, 2] = 0
    pts = pts[rng.permutation(len(pts))]
    pts = pts[: len(pts) - 500]  # some clusters lose points: sizes 1..3
    off, idx = check_extract(trk, oracle, pts, 0.4, 2, 3, brute=False)
    assert len(off) - 1 > 8192


def test_ihgp_c5(trk, oracle, synth):
    L, T = 40, 1000
    hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
    hyp_y = (np.exp(-5.0), np.exp(-3.0), np.exp(0.5))
    trk.ihgp_configure(0.1, 0.03, hyp, hyp_y, L)
    dt32 = float(np.float32(0.1))
    cx, cy = oracle.ihgp_setup(dt32, *hyp), oracle.ihgp_setup(dt32, *hyp_y)
    np.testing.assert_allclose(trk.ihgp_constants(0), cx, rtol=1e-12)
    np.testing.assert_allclose(trk.ihgp_constants(1), cy, rtol=1e-12)
    m_gpu, m_ref = np.zeros((T, 4)), np.zeros((T, 4))
    for frame in range(3):  # the smoothed state is carried across frames
        rings = synth.make_rings_c5(T, L, frame=frame)
        pv = trk.ihgp_step(rings, m_gpu)
        pv_ref = oracle.ihgp_step(rings, m_ref, 0.1, 0.03, cx, cy)
        np.testing.assert_allclose(pv, pv_ref, rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(m_gpu, m_ref, rtol=RTOL, atol=1e-9)
    assert np.abs(pv[:, 4:6]).max() <= 1.5


def test_error_paths(mot, trk):
    with pytest.raises(mot.MotError) as e:
        trk.set_cluster_params(0.0, 1, 10)
    assert e.value.code == -1
    bad = np.ones((10, 4), np.float32)
    bad[3, 1] = np.nan
    trk.set_cluster_params(0.3, 1, 10)
    with pytest.raises(mot.MotError) as e:
        trk.extract(bad)
    assert e.value.code == -6
    small = mot.Tracker(device=0, max_points=100, max_tracks=0)
    with pytest.raises(mot.MotError) as e:
        small.extract(np.ones((101, 4), np.float32))
    assert e.value.code == -3
    with pytest.raises(mot.MotError) as e:
        small.remove_static(np.ones((10, 4), np.float32))
    assert e.value.code == -4
    small.close()


@pytest.mark.parametrize("leaf", [(0.1, 0.1, 2.0), (0.05, 0.05, 1.0), (0.5, 0.5, 0.5)])
def test_voxel_grid(trk, oracle, synth, leaf):
    # pcl::VoxelGrid with the tracker's (L, L, 20 L) leaf, ahead of removeStatic (MOT.cpp:452-456)
    cloud, _ = synth.make_frame_c1()
    out = trk.voxel_grid(cloud, leaf)
    ref = oracle.voxel_grid(cloud, leaf)
    assert len(out) == len(ref) and 0 < len(out) < len(cloud)
    np.testing.assert_allclose(out[:, :3], ref[:, :3], rtol=RTOL, atol=1e-5)
    assert len(trk.voxel_grid(np.zeros((0, 4), np.float32), leaf)) == 0


def test_voxel_grid_then_frame(trk, oracle, synth):
    # the whole of clusterPointCloud (MOT.cpp:444-505): VoxelGrid -> removeStatic -> extract -> getCentroid
    occ, res, origin = synth.make_map_c1()
    cloud, _ = synth.make_frame_c1()
    L = 0.05
    down = trk.voxel_grid(cloud, (L, L, 20 * L))
    trk.set_map(occ, res, origin[:2], static_tolarance=2)
    trk.set_cluster_params(0.3, 5, 300)
    out = trk.frame(down, stamp_minus_time_init=1.0)
    kept_ref, _ = oracle.remove_static(down, occ, res, origin[:2], static_tolerance=2)
    off_ref, idx_ref = oracle.cluster_kdtree(kept_ref, 0.3, 5, 300)
    assert np.array_equal(out["kept"], kept_ref)
    assert np.array_equal(out["offsets"], off_ref) and np.array_equal(out["indices"], idx_ref)


@pytest.mark.parametrize("layout", ["xyz16", "xyzi_rgb32_unaligned", "bigendian"])
def test_pointcloud2_ingest(trk, oracle, layout):
    # sensor_msgs/PointCloud2 -> pcl::PointXYZ (pcl::fromROSMsg, MOT.cpp:448-449)
    rng = np.random.default_rng(9)
    n = 50000
    xyz = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
    xyz[rng.integers(0, n, 300), rng.integers(0, 3, 300)] = np.nan
    xyz[rng.integers(0, n, 50), 0] = np.inf
    if layout == "xyz16":
        step, offs, be = 16, (0, 4, 8), False
    elif layout == "xyzi_rgb32_unaligned":
        step, offs, be = 35, (1, 9, 22), False
    else:
        step, offs, be = 20, (0, 4, 8), True
    raw = rng.integers(0, 256, (n, step), dtype=np.uint8)
    for d, off in enumerate(offs):
        raw[:, off:off + 4] = xyz[:, d].astype(">f4" if be else "<f4").view(np.uint8).reshape(n, 4)
    for drop in (False, True):
        out = trk.unpack_pointcloud2(raw, n, step, offs, is_bigendian=be, drop_nonfinite=drop)
        ref = oracle.unpack_pointcloud2(raw, n, step, offs, is_bigendian=be, drop_nonfinite=drop)
        assert out.shape == ref.shape and np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert len(ref) < n


def test_obstacle_table(trk, oracle, synth):
    # the payload of publishObstacles (MOT.cpp:253-295), packed per track by the IHGP kernel
    T, L = 200, 10
    hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
    trk.ihgp_configure(0.1, 0.03, hyp, hyp, L)
    rings = synth.make_rings_c5(T, L)
    ids = np.arange(1000, 1000 + T, dtype=np.int32)[::-1].copy()
    m1, m2 = np.zeros((T, 4)), np.zeros((T, 4))
    pv, obs = trk.ihgp_step_obstacles(rings, m1, ids)
    pv2 = trk.ihgp_step(rings, m2)
    assert np.array_equal(pv, pv2) and np.array_equal(m1, m2)
    ref = oracle.obstacle_table(pv, ids)
    got = np.c_[obs["id"], obs["radius"], obs["x"], obs["y"], obs["vx"], obs["vy"], obs["vel_cov"]]
    assert np.array_equal(got, ref)


def _tracker_scenario(seed, n_frames, n_obj, dt=0.1):
    """Moving objects with drop-outs (gaps > 3 dt trigger the interpolation), births, deaths, two centroids close enough
    to hit the same track, and a long quiet stretch so that the 5 s purge fires."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-10, 10, (n_obj, 2))
    vel = rng.uniform(-1.2, 1.2, (n_obj, 2))
    born = rng.integers(0, n_frames // 3, n_obj)
    dies = born + rng.integers(n_frames // 4, n_frames, n_obj)
    frames = []
    for f in range(n_frames):
        t = 100.0 + f * dt
        cen = []
        for o in range(n_obj):
            if not (born[o] <= f < dies[o]):
                continue
            if rng.random() < 0.08 or (o % 5 == 0 and 30 <= f % 60 < 36):  # missed detections; periodic 6-frame gaps
                continue
            p = pos[o] + vel[o] * (f * dt) + rng.normal(0, 0.01, 2)
            cen.append([p[0], p[1], 0.0, t])
            if o % 7 == 0 and f % 11 == 0:
                cen.append([p[0] + 0.05, p[1] - 0.04, 0.0, t])  # a second centroid inside id_threshold of the same track
        if 70 <= f < 75:
            cen = []  # "No obstacles around"
        frames.append((t, np.array(cen, dtype=np.float32).reshape(-1, 4)))
    return frames


@pytest.mark.parametrize("seed,L,fast", [(1, 10, 1), (2, 40, 1), (2, 40, 0), (3, 7, 1), (4, 70, 1), (5, 33, 1)])
def test_tracks_association_lifecycle(mot, oracle, seed, L, fast):
    # cloudCallback's association / interpolation / registration / callIHGP / purge, device resident.
    # fast = 1: k_associate_fast (shared-memory table of last observations) + k_tracks_apply; 0: the one-kernel k_associate
    from oracle.tracker_ref import TrackerRef
    hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
    freq, thr = 10.0, 0.4
    saved = os.environ.get("MOT_ASSOC_FAST")
    os.environ["MOT_ASSOC_FAST"] = str(fast)
    try:
        t = mot.Tracker(device=0, max_points=1024, max_tracks=512)
    finally:
        if saved is None:
            os.environ.pop("MOT_ASSOC_FAST", None)
        else:
            os.environ["MOT_ASSOC_FAST"] = saved
    t.ihgp_configure(0.1, 0.03, hyp, hyp, L)
    ref = TrackerRef(freq, thr, L, 0.03, hyp, hyp)
    produced = 0
    for now, cen in _tracker_scenario(seed, 140, 24):
        out = t.tracks_step(cen, now, thr, freq)
        r = ref.step(cen, now)
        assert out["produced"] == (r is not None)
        assert out["n_tracks"] == len(ref.obj_ids)
        if r is not None:
            produced += 1
            ids_ref, pv_ref = r
            assert np.array_equal(out["ids"], ids_ref)
            np.testing.assert_allclose(out["pos_vel"], pv_ref, rtol=RTOL, atol=1e-6)
            assert np.array_equal(out["obstacles"]["id"], ids_ref)
        ids, rings, m = t.tracks_get()
        assert np.array_equal(ids, np.array(ref.obj_ids, dtype=np.int32))
        if len(ids):
            assert np.array_equal(rings.view(np.uint32), np.array(ref.stack, dtype=np.float32).view(np.uint32))  # rings bit-exact
            np.testing.assert_allclose(m, np.array(ref.m), rtol=RTOL, atol=1e-9)
    assert produced > 100 and ref.next_obj_num > 24  # births beyond the first frame happened, and the purge ran:
    assert len(ref.obj_ids) < ref.next_obj_num
    t.close()


def test_voxel_grid_index_overflow_passes_input_through(mot):
    # pcl::VoxelGrid::applyFilter warns and returns the input cloud when the voxel index would overflow an int
    t = mot.Tracker(device=0, max_points=4096, max_tracks=0)
    rng = np.random.default_rng(3)
    pts = np.ones((1000, 4), np.float32)
    pts[:, :3] = rng.uniform(-5000, 5000, (1000, 3))
    out = t.voxel_grid(pts, (0.001, 0.001, 0.001))
    assert t.last_warning == mot.MOT_WARN_VOXEL_OVERFLOW
    assert np.array_equal(out, pts)
    out = t.voxel_grid(pts, (2500.0, 2500.0, 2500.0))
    assert t.last_warning == 0 and 0 < len(out) <= 125
    t.close()


def test_tracks_table_full_is_not_fatal(mot, oracle):
    # the reference has no track cap; when max_tracks is reached the extra centroids are skipped (id -1, zero
    # rows, warning status) while filtering, the callback counter and the purge keep running -- so the table drains again
    from oracle.tracker_ref import TrackerRef
    hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
    freq, thr, L, T = 10.0, 0.4, 10, 8
    t = mot.Tracker(device=0, max_points=1024, max_tracks=T)
    t.ihgp_configure(0.1, 0.03, hyp, hyp, L)
    ref = TrackerRef(freq, thr, L, 0.03, hyp, hyp)

    def cen(n, now, x0=0.0):
        return np.array([[x0 + 3.0 * i, 1.0, 0.0, now] for i in range(n)], dtype=np.float32)

    now = 0.0
    out = t.tracks_step(cen(T, now), now, thr, freq)          # first frame: fills the table exactly
    ref.step(cen(T, now), now)
    assert t.last_warning == 0 and out["n_tracks"] == T
    for f in range(1, 4):                                      # 5 known objects + 3 new ones that cannot be registered
        now = 0.1 * f
        c = np.concatenate([cen(5, now), cen(3, now, x0=100.0)])
        out = t.tracks_step(c, now, thr, freq)
        assert t.last_warning == mot.MOT_WARN_TRACKS_FULL
        r_ids, r_pv = ref.step(cen(5, now), now)               # the reference run only sees what could be registered
        assert out["produced"] and out["n_tracks"] == T
        assert np.array_equal(out["ids"][:5], r_ids) and (out["ids"][5:] == -1).all()
        np.testing.assert_allclose(out["pos_vel"][:5], r_pv, rtol=RTOL, atol=1e-6)   # matched tracks were filtered as usual
        assert not out["pos_vel"][5:].any()
    # the old objects disappear.  The purge runs every 51st callback and drops tracks unseen for more than 5 s: the first one
    # (t = 5.1 s) still finds them 4.8 s old, the second one (t = 10.2 s) frees their slots and the two newcomers register
    for f in range(4, 4 + 110):
        now = 0.1 * f
        out = t.tracks_step(cen(2, now, x0=100.0), now, thr, freq)
    assert t.last_warning == 0 and out["n_tracks"] <= 3 and (out["ids"] >= 0).all()
    t.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("MOT_FUZZ_SEEDS", "12"))))
def test_fuzz_partition_vs_oracle(trk, oracle, seed):
    # randomised differential test: mixtures of dense blobs, planes, lines and uniform noise at random scales / tolerances;
    # exercises sparse and dense union-find tasks, 1-3 radix passes, empty / tiny clusters, min/max filters
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(200, 60000))
    scale = float(10 ** rng.uniform(-1, 2.3))
    tol = float(scale * 10 ** rng.uniform(-2.5, -0.7))
    parts = []
    left = n
    while left > 0:
        kind = rng.integers(0, 4)
        k = int(min(left, rng.integers(1, max(2, n // 3))))
        c = rng.uniform(-scale, scale, 3)
        if kind == 0:
            p = c + rng.normal(0, tol * rng.uniform(0.2, 3), (k, 3))
        elif kind == 1:
            p = c + np.c_[rng.uniform(-1, 1, (k, 2)) * scale * rng.uniform(0.01, 0.3), np.zeros(k)]
        elif kind == 2:
            p = c + np.outer(np.linspace(0, 1, k), rng.normal(0, 1, 3)) * scale * 0.3
        else:
            p = rng.uniform(-scale, scale, (k, 3))
        parts.append(p)
        left -= k
    pts = np.ones((n, 4), np.float32)
    pts[:, :3] = np.concatenate(parts)[rng.permutation(n)]
    mn = int(rng.integers(1, 8))
    mx = int(rng.choice([10, 300, 100000]))
    check_extract(trk, oracle, pts, tol, mn, mx, brute=n <= 4000)
