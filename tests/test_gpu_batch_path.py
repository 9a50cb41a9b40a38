"""GPU tests of the round-2 entry points: the full per-frame path on frame batches (removeStatic + clustering + tables +
circumcentres), packed 12-byte input, the fused clusterPointCloud call and mot_batch_run over several handles.
Each is compared with the oracle (bit-exact partition / removeStatic, rtol 1e-5 tables) and with the per-frame calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _frames_c1(synth, n_frames, n_points=20000):
    out = []
    for f in range(n_frames):
        cloud, _ = synth.make_frame_c1(n_points=n_points + 137 * f, frame=f, box_shift=(0.3 * f, -0.2 * f))
        out.append(cloud)
    return out


def _check_batch_against_oracle(res, clouds, oracle, occ, resn, origin, p, stamps, remove_static=True):
    fko, fco, off, idx = res["frame_kept_offsets"], res["frame_cluster_offsets"], res["offsets"], res["indices"]
    assert fco[0] == 0 and fco[-1] == res["K"] and off[0] == 0
    for f, cloud in enumerate(clouds):
        kept = oracle.remove_static(cloud, occ, resn, origin[:2], static_tolerance=p["static_tolerance"])[0] if remove_static else cloud
        assert fko[f + 1] - fko[f] == len(kept), f"frame {f}: kept count"
        o_ref, i_ref = oracle.cluster_kdtree(kept, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
        k0, k1 = fco[f], fco[f + 1]
        assert np.array_equal(off[k0:k1 + 1] - off[k0], o_ref), f"frame {f}: cluster offsets"
        assert np.array_equal(idx[off[k0]:off[k1]], i_ref), f"frame {f}: point indices"
        if res["centroids"] is not None and k1 > k0:
            cen_ref = oracle.get_centroid(kept, o_ref, i_ref, float(stamps[f]) if stamps is not None else 0.0)
            np.testing.assert_allclose(res["centroids"][k0:k1], cen_ref, rtol=RTOL, atol=1e-6)
        if res["stats"] is not None:
            for c in range(k1 - k0):
                pts = kept[i_ref[o_ref[c]:o_ref[c + 1]], :3]
                st = res["stats"][k0 + c]
                assert st["count"] == len(pts)
                np.testing.assert_allclose(st["mean"], pts.mean(0, dtype=np.float64), rtol=RTOL, atol=1e-5)
                assert np.array_equal(st["bbox_min"], pts.min(0)) and np.array_equal(st["bbox_max"], pts.max(0))


@pytest.mark.parametrize("packed12", [False, True])
def test_frame_batch_full_path(mot, oracle, synth, packed12):
    # removeStatic and circumcentres inside the batch call (clusterPointCloud, MOT.cpp:461-491, per frame)
    occ, resn, origin = synth.make_map_c1()
    p = synth.C1_PARAMS
    clouds = _frames_c1(synth, 5)
    clouds.insert(2, np.zeros((0, 4), np.float32))  # an empty frame in the middle
    stamps = np.arange(len(clouds), dtype=np.float32) * 0.1 + 1.0
    t = mot.Tracker(device=0, max_points=sum(len(c) for c in clouds) + 16, max_tracks=0)
    t.set_map(occ, resn, origin[:2], static_tolarance=p["static_tolerance"])
    t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    res = t.frame_batch(clouds, do_remove_static=True, stamps=stamps, packed12=packed12)
    _check_batch_against_oracle(res, clouds, oracle, occ, resn, origin, p, stamps)
    # and the per-frame call gives the same rows
    for f in (0, 3):
        one = t.frame(clouds[f], stamp_minus_time_init=float(stamps[f]))
        k0, k1 = res["frame_cluster_offsets"][f], res["frame_cluster_offsets"][f + 1]
        assert one["K"] == k1 - k0
        assert np.array_equal(one["centroids"].view(np.uint32), res["centroids"][k0:k1].view(np.uint32))
    # single-frame batch, without removeStatic
    res1 = t.frame_batch([clouds[1]], do_remove_static=False, stamps=None)
    _check_batch_against_oracle(res1, [clouds[1]], oracle, occ, resn, origin, p, None, remove_static=False)
    t.close()


def test_batch_run_over_several_handles(mot, oracle, synth):
    # SURVEY 8e / VERDICT r1 item 1: mot_batch_run shards the frames over handles (here three handles on one GPU -- the
    # code path is the same with one handle per GPU) and must leave exactly what one mot_frame_batch call leaves
    occ, resn, origin = synth.make_map_c1()
    p = synth.C1_PARAMS
    clouds = _frames_c1(synth, 7, n_points=12000)
    stamps = np.linspace(2.0, 2.6, len(clouds)).astype(np.float32)
    total = sum(len(c) for c in clouds)
    trks = [mot.Tracker(device=0, max_points=total + 16, max_tracks=0) for _ in range(3)]
    for t in trks:
        t.set_map(occ, resn, origin[:2], static_tolarance=p["static_tolerance"])
        t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    one = trks[0].frame_batch(clouds, do_remove_static=True, stamps=stamps)
    for n_h in (1, 2, 3):
        for packed12 in (False, True):
            many = mot.batch_run(trks[:n_h], clouds, do_remove_static=True, stamps=stamps, packed12=packed12)
            assert many["K"] == one["K"]
            for key in ("frame_kept_offsets", "frame_cluster_offsets", "offsets", "indices"):
                assert np.array_equal(many[key], one[key]), (n_h, key)
            assert np.array_equal(many["centroids"].view(np.uint32), one["centroids"].view(np.uint32))
            assert many["stats"].tobytes() == one["stats"].tobytes()
    _check_batch_against_oracle(one, clouds, oracle, occ, resn, origin, p, stamps)
    # more handles than frames: the spare handles get empty ranges
    two = mot.batch_run(trks, clouds[:2], do_remove_static=False, stamps=None)
    ref2 = trks[0].frame_batch(clouds[:2], do_remove_static=False)
    assert np.array_equal(two["offsets"], ref2["offsets"]) and np.array_equal(two["indices"], ref2["indices"])
    for t in trks:
        t.close()


@pytest.mark.parametrize("leaf", [0.0, 0.05])
def test_cluster_pointcloud2_fused(mot, oracle, synth, leaf):
    # fromROSMsg -> VoxelGrid -> removeStatic -> extract -> getCentroid in one call (MOT.cpp:448-491),
    # against the same stages called one by one and against the oracle
    occ, resn, origin = synth.make_map_c1()
    p = synth.C1_PARAMS
    cloud, _ = synth.make_frame_c1()
    rng = np.random.default_rng(5)
    n = len(cloud)
    xyz = cloud[:, :3].copy()
    xyz[rng.integers(0, n, 200), rng.integers(0, 3, 200)] = np.nan      # a non-dense cloud
    step, offs = 22, (2, 6, 14)                                         # unaligned records with other fields in between
    raw = rng.integers(0, 256, (n, step), dtype=np.uint8)
    for d, o in enumerate(offs):
        raw[:, o:o + 4] = xyz[:, d].astype("<f4").view(np.uint8).reshape(n, 4)
    t = mot.Tracker(device=0, max_points=n, max_tracks=0)
    t.set_map(occ, resn, origin[:2], static_tolarance=p["static_tolerance"])
    t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    out = t.cluster_pointcloud2(raw, n, step, offs, voxel_leaf_size=leaf, do_remove_static=True, stamp_minus_time_init=3.5)
    # staged: unpack (dropping non-finite points), VoxelGrid, then the frame call
    staged = t.unpack_pointcloud2(raw, n, step, offs, drop_nonfinite=True)
    if leaf > 0:
        staged = t.voxel_grid(staged, (leaf, leaf, 20 * leaf))
    fr = t.frame(staged, stamp_minus_time_init=3.5)
    assert out["m"] == fr["m"] and np.array_equal(out["kept"].view(np.uint32), fr["kept"].view(np.uint32))
    assert np.array_equal(out["offsets"], fr["offsets"]) and np.array_equal(out["indices"], fr["indices"])
    assert np.array_equal(out["centroids"].view(np.uint32), fr["centroids"].view(np.uint32))
    # oracle on the staged cloud
    kept_ref, _ = oracle.remove_static(staged, occ, resn, origin[:2], static_tolerance=p["static_tolerance"])
    off_ref, idx_ref = oracle.cluster_kdtree(kept_ref, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    assert np.array_equal(out["kept"], kept_ref) and np.array_equal(out["offsets"], off_ref) and np.array_equal(out["indices"], idx_ref)
    assert out["K"] > 5
    # without removeStatic, empty input
    out2 = t.cluster_pointcloud2(raw, n, step, offs, voxel_leaf_size=leaf, do_remove_static=False)
    assert out2["m"] == len(staged)
    out3 = t.cluster_pointcloud2(raw[:0], 0, step, offs)
    assert out3["m"] == 0 and out3["K"] == 0
    t.close()


def test_remove_static_two_pass_matches_one_pass(mot, oracle, synth):
    # MOT_RS_MODE=0 keeps the round-1 count + compact kernels selectable
    import os
    occ, resn, origin = synth.make_map_c1()
    cloud, _ = synth.make_frame_c1()
    os.environ["MOT_RS_MODE"] = "0"
    try:
        t0 = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
    finally:
        os.environ.pop("MOT_RS_MODE")
    t1 = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
    for t in (t0, t1):
        t.set_map(occ, resn, origin[:2], static_tolarance=2)
    a, b = t0.remove_static(cloud), t1.remove_static(cloud)
    ref, _ = oracle.remove_static(cloud, occ, resn, origin[:2], static_tolerance=2)
    assert np.array_equal(a, ref) and np.array_equal(b, ref)
    t0.close()
    t1.close()


def test_batch_run_across_gpus(mot, oracle, synth):
    # one handle per GPU (skipped on a single-GPU box): the merged tables, written to host memory, equal one handle's batch call
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least two GPUs")
    occ, resn, origin = synth.make_map_c1()
    p = synth.C1_PARAMS
    clouds = _frames_c1(synth, 9, n_points=15000)
    stamps = np.linspace(0.5, 1.3, len(clouds)).astype(np.float32)
    total = sum(len(c) for c in clouds)
    trks = [mot.Tracker(device=g, max_points=total + 16, max_tracks=0) for g in range(min(n_gpus, 4))]
    for t in trks:
        t.set_map(occ, resn, origin[:2], static_tolarance=p["static_tolerance"])
        t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    one = trks[0].frame_batch(clouds, do_remove_static=True, stamps=stamps)
    many = mot.batch_run(trks, clouds, do_remove_static=True, stamps=stamps, packed12=True)
    for key in ("frame_kept_offsets", "frame_cluster_offsets", "offsets", "indices"):
        assert np.array_equal(many[key], one[key]), key
    assert np.array_equal(many["centroids"].view(np.uint32), one["centroids"].view(np.uint32))
    assert many["stats"].tobytes() == one["stats"].tobytes()
    _check_batch_against_oracle(many, clouds, oracle, occ, resn, origin, p, stamps)
    for t in trks:
        t.close()


def _labels_parallel(oracle, clouds, tol, threads=8):
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(threads) as ex:   # the oracle is a C library: ctypes releases the GIL
        return list(ex.map(lambda c: oracle.labels_grid(c, tol), clouds))


def test_c3_full_shape(mot, oracle, synth):
    # BASELINE config 3 at the per-GPU shape of the 8-GPU run: 64 frames x 130k points in ONE batch call; every frame's
    # component labels against the oracle (full-size shapes under -m gpu)
    p3 = synth.C3_PARAMS
    sc3 = synth.scene_c3()
    nf = 64
    frames = [sc3.frame(f, n_points=synth.C3_POINTS) for f in range(nf)]
    t = mot.Tracker(device=0, max_points=nf * synth.C3_POINTS, max_tracks=0)
    t.set_cluster_params(p3["cluster_tolerance"], p3["min_cluster_size"], p3["max_cluster_size"])
    fco, off, idx = t.extract_batch(frames)
    lab = t.result_labels()
    refs = _labels_parallel(oracle, frames, p3["cluster_tolerance"])
    for f in range(nf):
        assert np.array_equal(lab[f * synth.C3_POINTS:(f + 1) * synth.C3_POINTS], refs[f] + f * synth.C3_POINTS), f"frame {f}"
    for f in (0, 31, 63):   # and the CSR of a few frames
        o_ref, i_ref = oracle.csr_from_labels(refs[f], p3["min_cluster_size"], p3["max_cluster_size"])
        k0, k1 = fco[f], fco[f + 1]
        assert np.array_equal(off[k0:k1 + 1] - off[k0], o_ref) and np.array_equal(idx[off[k0]:off[k1]], i_ref)
    t.close()


def test_c4_full_shape(mot, oracle, synth):
    # BASELINE config 4 at full size: 2^22 points, 2,000 blobs, tolerance 0.1 / 0.3 / 1.0 (sparse cells ... hundreds of points per
    # cell, heavy witness lists); labels and CSR against the oracle
    p4 = synth.C4_PARAMS
    f4 = synth.make_frame_c4()
    tols = (0.1, 0.3, 1.0)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(3) as ex:
        futs = {tol: ex.submit(oracle.labels_grid, f4, tol) for tol in tols}
        t = mot.Tracker(device=0, max_points=len(f4), max_tracks=0)
        got = {}
        for tol in tols:
            t.set_cluster_params(tol, p4["min_cluster_size"], p4["max_cluster_size"])
            off, idx = t.extract(f4)
            got[tol] = (t.result_labels(), off, idx)
        t.close()
        for tol in tols:
            ref = futs[tol].result()
            lab, off, idx = got[tol]
            assert np.array_equal(lab, ref), f"tol {tol}: labels"
            o_ref, i_ref = oracle.csr_from_labels(ref, p4["min_cluster_size"], p4["max_cluster_size"])
            assert np.array_equal(off, o_ref) and np.array_equal(idx, i_ref), f"tol {tol}: CSR"


def test_frame_batch_edge_cases(mot, oracle, synth):
    # everything removed by the map; empty frames first / last; a batch of only empty frames
    occ, resn, origin = synth.make_map_c1()
    p = synth.C1_PARAMS
    cloud, _ = synth.make_frame_c1(n_points=8000)
    t = mot.Tracker(device=0, max_points=1 << 16, max_tracks=0)
    t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    blocked = np.full_like(occ, 100)
    t.set_map(blocked, resn, origin[:2], static_tolarance=0)
    res = t.frame_batch([cloud, cloud[:100]], do_remove_static=True)
    assert res["K"] == 0 and not res["frame_kept_offsets"].any() and not res["frame_cluster_offsets"].any() and len(res["indices"]) == 0
    t.set_map(occ, resn, origin[:2], static_tolarance=p["static_tolerance"])
    empty = np.zeros((0, 4), np.float32)
    clouds = [empty, cloud, empty, empty, cloud[::2].copy(), empty]
    res = t.frame_batch(clouds, do_remove_static=True, stamps=np.arange(6, dtype=np.float32))
    _check_batch_against_oracle(res, clouds, oracle, occ, resn, origin, p, np.arange(6, dtype=np.float32))
    assert res["frame_cluster_offsets"][1] == 0 and res["frame_cluster_offsets"][2] == res["frame_cluster_offsets"][4]
    res = t.frame_batch([empty, empty], do_remove_static=True)
    assert res["K"] == 0 and len(res["indices"]) == 0 and not res["frame_kept_offsets"].any()
    t.close()


def _cluster_and_check(t, oracle, cloud, p):
    off, idx = t.extract(cloud)
    o_ref, i_ref = oracle.cluster_kdtree(cloud, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    assert np.array_equal(off, o_ref) and np.array_equal(idx, i_ref)
    return len(o_ref) - 1


@pytest.mark.parametrize("uf_mode", ["1", "2"])
def test_speculative_grid_plan_hits_and_misses(mot, oracle, synth, monkeypatch, uf_mode):
    """The handle reuses the grid of its previous call (mot_grid_plan): same extent -> hit, a cloud that grew or moved out of the
    padded grid -> miss and a second pass from its own bounding box, a different tolerance or frame count -> planned afresh.
    Every result equals the oracle's, whichever way it was planned."""
    monkeypatch.setenv("MOT_UF_MODE", uf_mode)
    p = dict(synth.C1_PARAMS)
    base, _ = synth.make_frame_c1(n_points=12000)
    t = mot.Tracker(device=0, max_points=1 << 16, max_tracks=0)
    t.small_frames(0)   # frames of this size would take the single-launch path, which plans nothing
    t.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    assert t.grid_plan() == (0, 0)
    k0 = _cluster_and_check(t, oracle, base, p)              # first call: planned from the bounding box
    assert k0 > 0 and t.grid_plan() == (0, 0)
    _cluster_and_check(t, oracle, base, p)                   # same cloud: hit
    assert t.grid_plan() == (1, 0)
    jitter = base.copy()
    jitter[:, :3] += np.float32(0.01) * np.random.default_rng(3).standard_normal((len(base), 3)).astype(np.float32)
    _cluster_and_check(t, oracle, jitter, p)                 # a slightly different frame of the same extent: hit
    assert t.grid_plan() == (2, 0)
    moved = base.copy()
    moved[:, 0] += np.float32(500.0)                         # the whole scene far outside the planned grid: miss
    _cluster_and_check(t, oracle, moved, p)
    assert t.grid_plan() == (2, 1)
    _cluster_and_check(t, oracle, moved, p)                  # the miss left a fresh plan behind
    assert t.grid_plan() == (3, 1)
    grown = np.concatenate([moved, moved[:50] + np.float32([0, 0, 300.0, 0])])   # one axis outgrows its bits: miss
    _cluster_and_check(t, oracle, grown, p)
    assert t.grid_plan() == (3, 2)
    one = grown.copy()
    one[7, 1] = np.float32(-1.0e4)                           # a single outlier is enough
    _cluster_and_check(t, oracle, one, p)
    assert t.grid_plan()[1] == 3
    # another tolerance drops the plan (no hit, no miss); a batch call has another frame count (the same)
    h, m = t.grid_plan()
    p2 = dict(p, cluster_tolerance=0.2)
    t.set_cluster_params(p2["cluster_tolerance"], p2["min_cluster_size"], p2["max_cluster_size"])
    _cluster_and_check(t, oracle, base, p2)
    assert t.grid_plan() == (h, m)
    _cluster_and_check(t, oracle, base, p2)
    assert t.grid_plan() == (h + 1, m)
    res = t.frame_batch([base, jitter], do_remove_static=False)
    assert t.grid_plan() == (h + 1, m)
    res2 = t.frame_batch([base, jitter], do_remove_static=False)
    assert t.grid_plan() == (h + 2, m)
    for k in ("offsets", "indices", "frame_cluster_offsets"):
        assert np.array_equal(res[k], res2[k])
    res3 = t.frame_batch([moved, jitter], do_remove_static=False)   # batch miss
    assert t.grid_plan() == (h + 2, m + 1)
    o_ref, i_ref = oracle.cluster_kdtree(moved, p2["cluster_tolerance"], p2["min_cluster_size"], p2["max_cluster_size"])
    k1 = res3["frame_cluster_offsets"][1]
    assert np.array_equal(res3["offsets"][:k1 + 1], o_ref) and np.array_equal(res3["indices"][:o_ref[-1]], i_ref)
    # NaN under a speculative plan is still the documented error, and the handle works afterwards
    bad = base.copy()
    bad[11, 2] = np.nan
    with pytest.raises(mot.MotError):
        t.extract(bad)
    _cluster_and_check(t, oracle, base, p2)
    # switched off: every call takes its own bounding box, same results
    t.grid_plan(False)
    h, m = t.grid_plan()
    _cluster_and_check(t, oracle, base, p2)
    _cluster_and_check(t, oracle, base, p2)
    assert t.grid_plan() == (h, m)
    t.close()
