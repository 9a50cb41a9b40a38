"""Generates tests/golden/ref_vectors.npz: input / output vectors produced by the REFERENCE's own code.

Run in the build container only (it compiles and runs the reference sources under /root/reference, which does not exist
on the GPU box):
    python tests/golden/make_ref_fixtures.py

The generator drives oracle/_ref/libmot_ref.so (the reference's ObstacleTrack / InfiniteHorizonGP / Matern32model sources
compiled against the stand-in headers of oracle/shim, see oracle/ref_harness.cpp) and records, for seeded inputs:

  ihgp_*      the stationary constants of InfiniteHorizonGP for several hyper-parameter sets; callIHGP outputs and GP
              means over consecutive calls (data_length 10 and 40)
  rs_*        removeStatic keep masks on the reference's own map fixture (map/sim_01) -- axis-aligned and rotated
              origin, static_tolarance 0 / 2 / 4 -- for points whose window stays inside the map
  gc_*        getCentroid outputs for seeded clusters (CSR), with a stamp / time_init offset
  trk_*       a 140-frame scenario through cloudCallback: the centroids clusterPointCloud produced, the ids and the
              position / velocity rows of every published ObstacleArrayMsg, the tracker lists after every fifth frame
The tests (tests/test_ref_pin.py) compare the CPU oracle and the CUDA path against these vectors.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
HYP_DEFAULT = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))  # launch-file values (MOT.cpp:106-112), exponentiated


def as_cloud(xyz):
    c = np.ones((len(xyz), 4), dtype=np.float32)
    c[:, :3] = xyz
    return c


def ihgp_vectors(out):
    hyps = np.array([HYP_DEFAULT, (np.exp(-4.0), np.exp(-2.0), np.exp(0.0)), (0.5, 1.0, 1.1), (1e-3, 0.2, 3.0), (1.0, 1.0, 1.0)])
    # dt_gp is a float member of the reference (MOT.h:112), so every dt is a float32 value
    dts = np.array([np.float32(0.1), np.float32(0.05), np.float32(0.125), np.float32(0.2), np.float32(1 / 15.0)], dtype=np.float64)
    out["ihgp_hyp"], out["ihgp_dt"] = hyps, dts
    out["ihgp_consts"] = np.array([ref.ihgp_constants(dt, *h) for dt, h in zip(dts, hyps)])
    for tag, L, T, steps in (("l10", 10, 32, 5), ("l40", 40, 8, 3)):
        rng = np.random.default_rng(40 + L)
        R = ref.Reference(frequency=10.0, lpf_tau=0.03, data_length=L)
        # rings: random walks with jumps big enough to hit the 1.5 m/s clamp on some tracks; stamps in .w
        vel = rng.uniform(-2.0, 2.0, (T, 1, 2))
        t = 100.0 + 0.1 * np.arange(L + steps)
        walk = rng.uniform(-10, 10, (T, 1, 2)) + vel * (t - 100.0)[None, :, None] + rng.normal(0, 0.02, (T, L + steps, 2))
        seq_rings, seq_pv, seq_m = [], [], []
        for s in range(steps):
            rings = np.zeros((T, L, 4), dtype=np.float32)
            rings[:, :, :2] = walk[:, s:s + L]
            rings[:, :, 3] = t[s:s + L]
            pv, m = R.call_ihgp(rings)
            seq_rings.append(rings), seq_pv.append(pv), seq_m.append(m)
        out[f"ihgp_{tag}_rings"], out[f"ihgp_{tag}_pos_vel"], out[f"ihgp_{tag}_m"] = np.array(seq_rings), np.array(seq_pv), np.array(seq_m)
        R.close()


def remove_static_vectors(out):
    g = np.load(os.path.join(GOLD, "sim_01_occupancy.npz"))
    occ, res, origin = g["occ"], np.float32(g["resolution"]), g["origin"]
    H, W = occ.shape
    rng = np.random.default_rng(7)
    cases = []
    for yaw, tol in ((0.0, 2), (0.0, 0), (0.0, 4), (0.3, 2), (-1.1, 1)):
        quat = np.array([0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2)])
        n = 2500
        # map-frame coordinates whose (2*tol+1)^2 window stays inside the grid, then rotated into the world frame
        u = rng.uniform((tol + 1) * res, (W - tol - 1) * res, n)
        v = rng.uniform((tol + 1) * res, (H - tol - 1) * res, n)
        x = origin[0] + np.cos(yaw) * u - np.sin(yaw) * v
        y = origin[1] + np.sin(yaw) * u + np.cos(yaw) * v
        pts = as_cloud(np.stack([x, y, rng.uniform(0, 2, n)], 1))
        R = ref.Reference(static_tolarance=tol)
        R.set_map(occ, res, origin[:2], quat)
        kept = R.remove_static(pts)
        R.close()
        keys = {p.tobytes() for p in kept}
        mask = np.array([p.tobytes() in keys for p in pts], dtype=np.uint8)
        assert mask.sum() == len(kept) and np.array_equal(pts[mask.astype(bool)], kept)  # order preserved
        cases.append((quat, tol, pts, mask))
        print(f"removeStatic yaw {yaw:+.1f} tol {tol}: kept {len(kept)} / {n}")
    out["rs_quat"] = np.array([c[0] for c in cases])
    out["rs_tol"] = np.array([c[1] for c in cases], dtype=np.int32)
    out["rs_pts"] = np.array([c[2] for c in cases])
    out["rs_keep"] = np.array([c[3] for c in cases])


def get_centroid_vectors(out):
    rng = np.random.default_rng(11)
    n, K = 3000, 80
    pts = as_cloud(np.c_[rng.uniform(-20, 20, (n, 2)), rng.uniform(0, 2, n)])
    # clusters: blobs of 3..60 points around random centres, index order shuffled as a BFS would leave it
    sizes = rng.integers(3, 61, K)
    perm = rng.permutation(n)[: sizes.sum()]
    off = np.zeros(K + 1, dtype=np.int32)
    off[1:] = np.cumsum(sizes)
    for c in range(K):
        seg = perm[off[c]:off[c + 1]]
        centre = rng.uniform(-15, 15, 2)
        shape = rng.integers(0, 3)
        if shape == 0:      # blob
            xy = centre + rng.normal(0, 0.2, (len(seg), 2))
        elif shape == 1:    # thin arc (LiDAR return of a round object)
            a = rng.uniform(0, np.pi, len(seg))
            xy = centre + 0.3 * np.stack([np.cos(a), np.sin(a)], 1) + rng.normal(0, 0.005, (len(seg), 2))
        else:               # L-shaped corner
            s = rng.uniform(0, 1, len(seg))
            xy = centre + np.where((rng.random(len(seg)) < 0.5)[:, None], np.stack([s, 0 * s], 1), np.stack([0 * s, s], 1))
        pts[seg, :2] = xy.astype(np.float32)
    idx = perm.astype(np.int32)
    R = ref.Reference()
    stamp, time_init = 1234.5, 1000.25
    cen = R.get_centroid(pts, off, idx, stamp=stamp, time_init=time_init)
    R.close()
    out["gc_pts"], out["gc_off"], out["gc_idx"], out["gc_stamp"], out["gc_centroids"] = pts, off, idx, np.array([stamp, time_init]), cen
    print("getCentroid:", K, "clusters")


def tracker_scenario(seed, n_frames=140, n_obj=16, dt=0.1):
    """Small moving objects (rings of 7-10 returns, radius 0.1 m) with drop-outs, gaps > 3 dt, births, deaths, a companion
    0.33 m away (inside id_threshold, outside the cluster tolerance) and a stretch without returns."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-8, 8, (n_obj, 2))
    vel = rng.uniform(-1.0, 1.0, (n_obj, 2))
    born = rng.integers(0, n_frames // 3, n_obj)
    dies = born + rng.integers(n_frames // 4, n_frames, n_obj)
    ang = np.linspace(0, 2 * np.pi, 10, endpoint=False)
    ring = np.stack([0.1 * np.cos(ang), 0.1 * np.sin(ang)], 1)
    frames = []
    for f in range(n_frames):
        t = 100.0 + f * dt
        parts = []
        for o in range(n_obj):
            if not (born[o] <= f < dies[o]):
                continue
            if rng.random() < 0.08 or (o % 5 == 0 and 30 <= f % 60 < 36):
                continue
            p = pos[o] + vel[o] * (f * dt) + rng.normal(0, 0.01, 2)
            k = 10 if o % 3 else 7
            parts.append(np.c_[ring[:k] + p + rng.normal(0, 0.003, (k, 2)), rng.uniform(0.2, 0.8, k)])
            if o % 7 == 0 and f % 11 == 0:
                parts.append(np.c_[ring + p + np.array([0.33, 0.0]), np.full(10, 0.5)])
        if 70 <= f < 75 or not parts:
            parts = [np.zeros((0, 3))]
        frames.append((t, as_cloud(np.concatenate(parts))))
    return frames


TRK_PARAMS = dict(frequency=10.0, cluster_tolerance=0.15, min_cluster_size=5, max_cluster_size=200, voxel_leaf_size=0.05,
                  static_tolarance=2, id_threshold=0.4, lpf_tau=0.03, data_length=10)


def tracker_vectors(out):
    occ = np.zeros((1400, 1400), dtype=np.int8)  # 70 m x 70 m of free space: nothing is static
    res, origin = np.float32(0.05), (-35.0, -35.0)
    frames = tracker_scenario(1)
    R = ref.Reference(**TRK_PARAMS)
    R.set_map(occ, res, origin)
    pts, pt_off, stamps = [], [0], []
    cen, cen_off = [], [0]
    ids, pv, out_off, produced = [], [], [0], []
    t_ids, t_rings, t_m, t_off, t_frames = [], [], [], [0], []
    for f, (t, cloud) in enumerate(frames):
        c = R.cluster_point_cloud(cloud, t)   # what cloudCallback is about to see (no tracker state is touched)
        r = R.cloud_callback(cloud, t)
        pts.append(cloud), pt_off.append(pt_off[-1] + len(cloud)), stamps.append(t)
        cen.append(c), cen_off.append(cen_off[-1] + len(c))
        produced.append(r is not None)
        if r is not None:
            ids.append(r[0]), pv.append(r[1])
        out_off.append(out_off[-1] + (len(r[0]) if r is not None else 0))
        a, b, m = R.tracks()
        if f % 5 == 0 or f == len(frames) - 1:   # the tracker lists, every fifth frame and at the end
            t_ids.append(a), t_rings.append(b), t_m.append(m), t_off.append(t_off[-1] + len(a)), t_frames.append(f)
    print("tracker: frames", len(frames), "produced", int(np.sum(produced)), "ids handed out", R.next_obj_num(), "alive at the end", len(a))
    R.close()
    out["trk_params"] = np.array([TRK_PARAMS[k] for k in sorted(TRK_PARAMS)], dtype=np.float64)
    out["trk_param_names"] = np.array(sorted(TRK_PARAMS))
    out["trk_map"] = np.array([1400, 1400, 0.05, -35.0, -35.0])
    out["trk_pts"], out["trk_pt_off"], out["trk_stamps"] = np.concatenate(pts), np.array(pt_off, dtype=np.int64), np.array(stamps)
    out["trk_centroids"], out["trk_cen_off"] = np.concatenate(cen), np.array(cen_off, dtype=np.int64)
    out["trk_produced"] = np.array(produced, dtype=np.uint8)
    out["trk_ids"], out["trk_pos_vel"], out["trk_out_off"] = np.concatenate(ids), np.concatenate(pv), np.array(out_off, dtype=np.int64)
    out["trk_track_ids"], out["trk_track_rings"], out["trk_track_m"] = np.concatenate(t_ids), np.concatenate(t_rings), np.concatenate(t_m)
    out["trk_track_off"], out["trk_track_frames"] = np.array(t_off, dtype=np.int64), np.array(t_frames, dtype=np.int64)


if __name__ == "__main__":
    assert ref.can_build(), "needs /root/reference (build container only)"
    ref.build()
    vec = {}
    ihgp_vectors(vec)
    remove_static_vectors(vec)
    get_centroid_vectors(vec)
    tracker_vectors(vec)
    path = os.path.join(GOLD, "ref_vectors.npz")
    np.savez_compressed(path, **vec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(vec), "arrays")
