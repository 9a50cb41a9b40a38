"""The BASELINE.json configs other than the headline one, run by bench.py on rank 0 at N = 1 (the `configs` block of its JSON
line): c1 (64k-point frame + map, the whole per-frame path), c3 (64 x 130k-point frames per GPU in one batch call),
c4 (2^22 points, 2,000 blobs, tolerance 0.1 / 0.3 / 1.0) and c5 (tracker step: frame + IHGP for 1,000 tracks).

Every entry carries: value (Mpoints/s, device-resident input, median of CUDA-event timings of the public call), ms, the
fraction of the HBM roofline for SURVEY 8d's frame model  B = 16 N + (152 + 16 P) M + 16 Cc + 44 K  bytes, and `parity`:
the result of the same call compared with the CPU oracle IN THIS RUN (bit-exact partition / removeStatic; rtol 1e-5 for
circumcentres and IHGP).  The oracle is only the checker here.  `kernels` adds roofline entries of the stage kernels the
headline workload does not exercise (removeStatic compaction, farthest pair, IHGP).
"""
import threading
import time

import numpy as np

UNIT = "Mpoints/s"


def _median_ms(trk, fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        trk.timer_start()
        fn()
        ts.append(trk.timer_stop())
    return float(np.median(ts))


def _frame_bytes(N, M, cc, K, key_bits):
    P = (key_bits + 9) // 10
    return 16 * N + (152 + 16 * P) * M + 16 * cc + 44 * K


def _entry(points, ms, alg_bytes, peak, parity, **extra):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    d = {"value": round(points / (ms * 1e-3) / 1e6, 1), "unit": UNIT, "ms": round(ms, 4), "points": int(points), "alg_bytes": int(alg_bytes),
         "gbs": round(gbs, 1), "frac": round(gbs / peak, 4), "parity": bool(parity)}
    d.update(extra)
    return d


def _kernel_rows(prof, reps, names, bytes_of, peak):
    rows = []
    for n in names:
        if n in prof:
            ms, cnt = prof[n]
            avg = ms / cnt
            b = bytes_of.get(n)
            rows.append({"kernel": n, "avg_us": round(avg * 1e3, 2), "launches_per_call": cnt / reps, "alg_bytes": b,
                         "gbs": round(b / (avg * 1e-3) / 1e9, 1) if b else None, "frac": round(b / (avg * 1e-3) / 1e9 / peak, 4) if b else None})
    return rows


def run(mot, oracle, device, peak, quick=False):
    import torch

    synth = mot.synth
    dev = torch.device("cuda", device)
    out = {}
    kernels = []

    # ---- oracle labels of the c4 frame run on host threads while the GPU part goes on (41 s for tol 1.0 on one core) ----
    p4 = synth.C4_PARAMS
    n4 = 1 << 20 if quick else 1 << 22
    f4 = synth.make_frame_c4(n_points=n4, n_blobs=500 if quick else 2000)
    c4_tols = (0.1, 0.3, 1.0)
    c4_ref = {}

    def c4_oracle(tol):
        c4_ref[tol] = oracle.labels_grid(f4, tol)

    c4_threads = [threading.Thread(target=c4_oracle, args=(tol,)) for tol in c4_tols]
    for th in c4_threads:
        th.start()

    # ---- c1: removeStatic + clustering + tables + circumcentres on one 64k-point frame ----------------------------------
    occ, res, origin = synth.make_map_c1()
    cloud, _ = synth.make_frame_c1()
    p = synth.C1_PARAMS
    trk = mot.Tracker(device=device, max_points=max(n4, 1 << 20), max_tracks=2048)
    trk.set_map(occ, res, origin[:2], static_tolarance=p["static_tolerance"])
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    o1 = trk.frame(cloud, 1.0)
    kept_ref, _ = oracle.remove_static(cloud, occ, res, origin[:2], static_tolerance=p["static_tolerance"])
    off_ref, idx_ref = oracle.cluster_kdtree(kept_ref, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    cen_ref = oracle.get_centroid(kept_ref, off_ref, idx_ref, 1.0)
    ok1 = (np.array_equal(o1["kept"], kept_ref) and np.array_equal(o1["offsets"], off_ref) and np.array_equal(o1["indices"], idx_ref)
           and np.allclose(o1["centroids"], cen_ref, rtol=1e-5, atol=1e-6))
    d_cloud = torch.from_numpy(cloud).to(dev)
    ms_dev = _median_ms(trk, lambda: trk.frame_device(d_cloud.data_ptr(), len(cloud), True, True, 1.0), reps=15)
    launches = trk.last_launches()
    bufs = trk.frame_buffers(len(cloud))  # caller-owned output buffers, as a C++ caller of mot_frame keeps them (pageable memory)
    ms_host = _median_ms(trk, lambda: trk.frame_into(cloud, bufs, 1.0), reps=15, warm=3)
    g = trk.result_grid()
    ph = trk.small_frame_phases() if launches <= 5 else {}
    dev_us = round(sum(v for k, v in ph.items() if k[:2].isdigit()) / 1e3, 1) if ph else None
    b1 = _frame_bytes(len(cloud), o1["m"], g["coarse_cells"], o1["K"], g["key_bits"])
    out["c1"] = _entry(len(cloud), ms_dev, b1, peak, ok1, call="mot_frame_device (removeStatic + clustering + tables + circumcentres)",
                       kept=int(o1["m"]), clusters=int(o1["K"]), launches=launches, host_buffers_ms=round(ms_host, 4),
                       roofline_us=round(b1 / (peak * 1e9) * 1e6, 2),
                       path="small-frame path: one CUDA graph of 5 kernels, one host round trip" if launches <= 5 else "general path",
                       device_us=dev_us, timing="ms: CUDA events around the call (launch + kernels + the wait); device_us: first to last kernel "
                       "instruction (%globaltimer stamps); host_buffers_ms: mot_frame with pageable caller-owned buffers, copies included")
    trk.set_profiling(True)
    reps = 5
    for _ in range(reps):
        trk.frame_device(d_cloud.data_ptr(), len(cloud), True, True, 1.0)
    prof = trk.profile()
    trk.set_profiling(False)
    kernels += _kernel_rows(prof, reps, ("k_compact_onepass<map>", "k_farthest_pair", "k_circumcentre", "k_fs_front", "k_fs_edges", "k_fs_tables",
                                         "k_fs_farthest", "k_fs_finish"),
                            {"k_compact_onepass<map>": 16 * len(cloud) + 16 * o1["m"], "k_circumcentre": 16 * int(off_ref[-1]) + 16 * o1["K"],
                             "k_fs_front": 16 * len(cloud) + 100 * o1["m"]}, peak)

    # ---- c5: tracker step = the c1 frame through the host-buffer call + IHGP for 1,000 tracks (L = 40) ----------------------
    T, L = 1000, 40
    hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
    trk.ihgp_configure(0.1, 0.03, hyp, hyp, L)
    rings = synth.make_rings_c5(T, L)
    m_gpu, m_ref = np.zeros((T, 4)), np.zeros((T, 4))
    cx = oracle.ihgp_setup(float(np.float32(0.1)), *hyp)
    pv = trk.ihgp_step(rings, m_gpu)
    pv_ref = oracle.ihgp_step(rings, m_ref, 0.1, 0.03, cx, cx)
    ok5 = ok1 and np.allclose(pv, pv_ref, rtol=1e-5, atol=1e-6) and np.allclose(m_gpu, m_ref, rtol=1e-5, atol=1e-9)

    def step_c5():
        trk.frame_into(cloud, bufs, 1.0)
        trk.ihgp_step(rings, m_gpu)

    ms5 = _median_ms(trk, step_c5, reps=9)
    ms_ihgp = _median_ms(trk, lambda: trk.ihgp_step(rings, m_gpu), reps=9)
    b5 = b1 + T * (16 * L + 96)
    # association + lifecycle + IHGP on the device-resident track table (mot_tracks_step, MOT.cpp:176-233) for T slowly moving objects
    # 4 m apart that arrive in a different order every frame; the ids must stay the ones of the first frame
    trk_t = mot.Tracker(device=device, max_points=1024, max_tracks=2 * T)
    trk_t.ihgp_configure(0.1, 0.03, hyp, hyp, L)
    rng5 = np.random.default_rng(5)
    side = int(np.ceil(np.sqrt(T)))
    base5 = np.stack(np.meshgrid(np.arange(side), np.arange(side)), -1).reshape(-1, 2)[:T].astype(np.float64) * 4.0
    vel5 = rng5.uniform(-0.5, 0.5, (T, 2))
    ts5, ok_ids = [], True
    for f in range(14):
        now = 0.1 * f
        cen5 = np.zeros((T, 4), np.float32)
        cen5[:, :2] = base5 + vel5 * now
        cen5[:, 3] = now
        order = np.arange(T) if f == 0 else rng5.permutation(T)  # registration order = object index
        t0 = time.perf_counter()
        o5 = trk_t.tracks_step(cen5[order], now, 1.0, 10.0)
        ts5.append((time.perf_counter() - t0) * 1e3)
        if f > 0:
            ok_ids = ok_ids and bool(np.array_equal(o5["ids"], order)) and o5["n_tracks"] == T
    trk_t.close()
    ms_tracks = float(np.median(ts5[4:]))
    ok5 = ok5 and ok_ids
    out["c5"] = _entry(len(cloud), ms5, b5, peak, ok5, call="mot_frame (host buffers) + mot_ihgp_step, 1000 tracks, L = 40", tracks=T,
                       ihgp_only_ms=round(ms_ihgp, 4), tracks_per_s=round(T / (ms_ihgp * 1e-3), 0), tracks_step_ms=round(ms_tracks, 4),
                       tracks_step="mot_tracks_step: association + lifecycle + IHGP for 1000 centroids against the device-resident table (host wall clock, "
                       "pageable buffers); ids checked against the first frame's")
    trk.set_profiling(True)
    for _ in range(reps):
        trk.ihgp_step(rings, m_gpu)
    prof = trk.profile()
    trk.set_profiling(False)
    kernels += _kernel_rows(prof, reps, ("k_ihgp_step",), {"k_ihgp_step": T * (16 * L + 96)}, peak)

    # ---- c3: 64 frames of 130k points in one batch call ---------------------------------------------------------------------
    p3 = synth.C3_PARAMS
    sc3 = synth.scene_c3()
    nf = 16 if quick else 64
    frames = [sc3.frame(f, n_points=synth.C3_POINTS) for f in range(nf)]
    allp = np.ascontiguousarray(np.concatenate(frames))
    fo = np.arange(nf + 1, dtype=np.int64) * synth.C3_POINTS
    d_all = torch.from_numpy(allp).to(dev)
    big = mot.Tracker(device=device, max_points=len(allp), max_tracks=0)
    big.set_cluster_params(p3["cluster_tolerance"], p3["min_cluster_size"], p3["max_cluster_size"])
    big.cluster_batch_device(d_all.data_ptr(), fo)
    lab = big.result_labels()
    ok3 = [True]

    def c3_check(f0, f1):
        for f in range(f0, f1):
            ref = oracle.labels_grid(frames[f], p3["cluster_tolerance"]) + f * synth.C3_POINTS
            if not np.array_equal(lab[f * synth.C3_POINTS:(f + 1) * synth.C3_POINTS], ref):
                ok3[0] = False

    ths = [threading.Thread(target=c3_check, args=(nf * i // 8, nf * (i + 1) // 8)) for i in range(8)]
    for th in ths:
        th.start()
    ms3 = _median_ms(big, lambda: big.cluster_batch_device(d_all.data_ptr(), fo), reps=5)
    M3, K3, _ = big.result_counts()
    g3 = big.result_grid()
    for th in ths:
        th.join()
    out["c3"] = _entry(len(allp), ms3, _frame_bytes(0, M3, g3["coarse_cells"], K3, g3["key_bits"]), peak, ok3[0],
                       call=f"mot_cluster_batch_device, {nf} frames x {synth.C3_POINTS} points", frames=nf, clusters=int(K3),
                       parity_frames_checked=nf)
    big.close()
    del d_all

    # ---- c4: dense 2^22-point frame, 2,000 blobs, tolerance sweep ----------------------------------------------------------
    d_f4 = torch.from_numpy(f4).to(dev)
    c4_rows = {}
    c4_labels = {}
    for tol in c4_tols:
        trk.set_cluster_params(tol, p4["min_cluster_size"], p4["max_cluster_size"])
        trk.frame_device(d_f4.data_ptr(), len(f4))
        c4_labels[tol] = trk.result_labels()
        ms4 = _median_ms(trk, lambda: trk.frame_device(d_f4.data_ptr(), len(f4)), reps=5, warm=1)
        M4, K4, _ = trk.result_counts()
        g4 = trk.result_grid()
        cnt = trk.result_counters()
        c4_rows[tol] = (ms4, _frame_bytes(0, M4, g4["coarse_cells"], K4, g4["key_bits"]), K4, g4, int(cnt[9]), int(cnt[10]))
    for th in c4_threads:
        th.join()
    for tol in c4_tols:
        ms4, b4, K4, g4, h1, h2 = c4_rows[tol]
        out[f"c4_tol{tol}"] = _entry(len(f4), ms4, b4, peak, np.array_equal(c4_labels[tol], c4_ref[tol]), call="mot_frame_device", tolerance=tol,
                                     clusters=int(K4), fine_cells=g4["fine_cells"], key_bits=g4["key_bits"], heavy_pairs=[h1, h2])
    trk.close()
    out["kernels"] = kernels
    out["all_parity_ok"] = all(v["parity"] for k, v in out.items() if isinstance(v, dict) and "parity" in v)
    return out
