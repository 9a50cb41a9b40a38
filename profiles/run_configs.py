"""Runs the five BASELINE.json configs on one B200: throughput of the GPU path and parity against the CPU oracle.

    python profiles/run_configs.py [--quick] > profiles/r01_configs.md

c1  64k-point frame + map, removeStatic + clustering + tables (mot_frame), launch-file min/max (5 / 300)
c2  one 2^20-point frame, tol 0.5, min 5 / max 100000 (and max 300 as in the launch file)
c3  batch of 130k-point frames (mot_cluster_batch_device); 64 frames per call
c4  2^22-point frame, 2,000 blobs, tolerance sweep 0.1 .. 1.0
c5  end-to-end tracker step: c1 frame (removeStatic + clustering + circumcentres) + IHGP for 1,000 tracks, L = 40
Parity = component labels of every point equal to the oracle's (orc_labels_grid: exact fp32 predicate on a CPU grid)
and, where the oracle finishes quickly, the full CSR against the restated PCL path (KD-tree + BFS).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

mot = entry.load_package()
oracle = entry.load_oracle()
synth = mot.synth
quick = "--quick" in sys.argv
rows = []


def gpu_time(fn, reps=5):
    fn()
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def add(cfg, what, n, secs, parity, note=""):
    rows.append((cfg, what, n, secs * 1e3, n / secs / 1e6, parity, note))
    print(f"| {cfg} | {what} | {n} | {secs * 1e3:.3f} | {n / secs / 1e6:.1f} | {parity} | {note} |", flush=True)


print("| config | call | points | best ms | Mpoints/s | parity vs oracle | note |")
print("|---|---|---|---|---|---|---|")

# ---- c1 -------------------------------------------------------------------------------------------------------------
occ, res, origin = synth.make_map_c1()
cloud, _ = synth.make_frame_c1()
p = synth.C1_PARAMS
trk = mot.Tracker(device=0, max_points=1 << 22, max_tracks=2048)
trk.set_map(occ, res, origin[:2], static_tolarance=p["static_tolerance"])
trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
out = trk.frame(cloud, 1.0)
kept_ref, _ = oracle.remove_static(cloud, occ, res, origin[:2], static_tolerance=p["static_tolerance"])
off_ref, idx_ref = oracle.cluster_kdtree(kept_ref, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
ok = np.array_equal(out["kept"], kept_ref) and np.array_equal(out["offsets"], off_ref) and np.array_equal(out["indices"], idx_ref)
d_cloud = torch.from_numpy(cloud).cuda()
t = gpu_time(lambda: trk.frame_device(d_cloud.data_ptr(), len(cloud), True, True, 1.0))
add("c1", "mot_frame_device (removeStatic+cluster+tables+circumcentre)", len(cloud), t, "bit-exact" if ok else "MISMATCH",
    f"kept {out['m']}, {out['K']} clusters")
t = gpu_time(lambda: trk.frame(cloud, 1.0))
add("c1", "mot_frame (host buffers)", len(cloud), t, "bit-exact" if ok else "MISMATCH", "incl. H2D/D2H")
t0 = time.perf_counter()
oracle.remove_static(cloud, occ, res, origin[:2], static_tolerance=2)
oracle.cluster_kdtree(kept_ref, 0.3, 5, 300)
oracle.get_centroid(kept_ref, off_ref, idx_ref, 1.0)
add("c1", "CPU oracle (1 thread)", len(cloud), time.perf_counter() - t0, "-", "restated reference path")

# ---- c2 -------------------------------------------------------------------------------------------------------------
p2 = synth.C2_PARAMS
frame = synth.scene_c2().frame(0)
d_frame = torch.from_numpy(frame).cuda()
lab_ref = oracle.labels_grid(frame, p2["cluster_tolerance"])
for mx in (p2["max_cluster_size"], 300):
    trk.set_cluster_params(p2["cluster_tolerance"], p2["min_cluster_size"], mx)
    trk.frame_device(d_frame.data_ptr(), len(frame))
    ok = np.array_equal(trk.result_labels(), lab_ref)
    r = trk.result_fetch()
    o_ref, i_ref = oracle.csr_from_labels(lab_ref, p2["min_cluster_size"], mx)
    ok = ok and np.array_equal(r["offsets"], o_ref) and np.array_equal(r["indices"], i_ref)
    t = gpu_time(lambda: trk.frame_device(d_frame.data_ptr(), len(frame)))
    add("c2", f"mot_frame_device, max_cluster_size {mx}", len(frame), t, "bit-exact" if ok else "MISMATCH", f"{r['K']} clusters")

# ---- c3 -------------------------------------------------------------------------------------------------------------
p3 = synth.C3_PARAMS
sc3 = synth.scene_c3()
nf = 16 if quick else 64
frames = [sc3.frame(f, n_points=synth.C3_POINTS) for f in range(nf)]
allp = np.ascontiguousarray(np.concatenate(frames))
fo = np.arange(nf + 1, dtype=np.int64) * synth.C3_POINTS
d_all = torch.from_numpy(allp).cuda()
big = mot.Tracker(device=0, max_points=len(allp), max_tracks=0)
big.set_cluster_params(p3["cluster_tolerance"], p3["min_cluster_size"], p3["max_cluster_size"])
big.cluster_batch_device(d_all.data_ptr(), fo)
lab = big.result_labels()
ok = True
for f in range(0, nf, max(1, nf // 8)):
    ref = oracle.labels_grid(frames[f], p3["cluster_tolerance"]) + f * synth.C3_POINTS
    ok = ok and np.array_equal(lab[f * synth.C3_POINTS:(f + 1) * synth.C3_POINTS], ref)
t = gpu_time(lambda: big.cluster_batch_device(d_all.data_ptr(), fo), reps=3)
add("c3", f"mot_cluster_batch_device, {nf} frames x 130k", len(allp), t, "bit-exact (8 frames checked)" if ok else "MISMATCH",
    f"{big.result_counts()[1]} clusters")
big.close()

# ---- c4 -------------------------------------------------------------------------------------------------------------
p4 = synth.C4_PARAMS
n4 = 1 << 20 if quick else 1 << 22
f4 = synth.make_frame_c4(n_points=n4, n_blobs=500 if quick else 2000)
d_f4 = torch.from_numpy(f4).cuda()
for tol in synth.C4_TOLERANCES:
    trk.set_cluster_params(tol, p4["min_cluster_size"], p4["max_cluster_size"])
    trk.frame_device(d_f4.data_ptr(), len(f4))
    check = tol in (0.1, 0.3, 1.0) or quick
    par = "not checked at this size"
    if check:
        ref = oracle.labels_grid(f4, tol)
        par = "bit-exact" if np.array_equal(trk.result_labels(), ref) else "MISMATCH"
    t = gpu_time(lambda: trk.frame_device(d_f4.data_ptr(), len(f4)), reps=3)
    g = trk.result_grid()
    add("c4", f"mot_frame_device, tol {tol}", len(f4), t, par, f"{trk.result_counts()[1]} clusters, {g['fine_cells']} fine cells, key {g['key_bits']} bits")

# ---- c5 -------------------------------------------------------------------------------------------------------------
T, L = 1000, 40
hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
trk.set_map(occ, res, origin[:2], static_tolarance=2)
trk.set_cluster_params(0.3, 5, 300)
trk.ihgp_configure(0.1, 0.03, hyp, hyp, L)
rings = synth.make_rings_c5(T, L)
m_gpu, m_ref = np.zeros((T, 4)), np.zeros((T, 4))
cx = oracle.ihgp_setup(float(np.float32(0.1)), *hyp)
pv = trk.ihgp_step(rings, m_gpu)
pv_ref = oracle.ihgp_step(rings, m_ref, 0.1, 0.03, cx, cx)
ok = np.allclose(pv, pv_ref, rtol=1e-5, atol=1e-6) and np.allclose(m_gpu, m_ref, rtol=1e-5, atol=1e-9)


def step_c5():
    trk.frame(cloud, 1.0)
    trk.ihgp_step(rings, m_gpu)


t = gpu_time(step_c5)
add("c5", "mot_frame + mot_ihgp_step (1,000 tracks, L=40), host buffers", len(cloud), t, "rtol 1e-5" if ok else "MISMATCH", "end-to-end tracker step")
t = gpu_time(lambda: trk.ihgp_step(rings, m_gpu))
add("c5", "mot_ihgp_step alone (tracks/s in the Mpoints/s column = Mtracks/s)", T, t, "rtol 1e-5" if ok else "MISMATCH", "incl. H2D/D2H of rings and state")
t0 = time.perf_counter()
oracle.ihgp_step(rings, m_ref, 0.1, 0.03, cx, cx)
add("c5", "CPU oracle IHGP (1 thread)", T, time.perf_counter() - t0, "-", "")
trk.close()
