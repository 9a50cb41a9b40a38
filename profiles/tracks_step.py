"""Wall clock of mot_tracks_step (association + lifecycle + IHGP, MOT.cpp:176-233) for T slowly moving centroids:
python profiles/tracks_step.py [T] [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as entry

mot = entry.load_package()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rng = np.random.default_rng(5)
side = int(np.ceil(np.sqrt(T)))
base = np.stack(np.meshgrid(np.arange(side), np.arange(side)), -1).reshape(-1, 2)[:T].astype(np.float64) * 4.0   # 4 m apart
vel = rng.uniform(-1, 1, (T, 2))
trk = mot.Tracker(device=0, max_points=1024, max_tracks=max(2 * T, 64))
hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
trk.ihgp_configure(0.1, 0.03, hyp, hyp, 40)
ts = []
for f in range(frames):
    now = 0.1 * f
    cen = np.zeros((T, 4), np.float32)
    cen[:, :2] = base + vel * now
    cen[:, 3] = now
    order = rng.permutation(T)   # clusters arrive in size order, not in track order
    t0 = time.perf_counter()
    out = trk.tracks_step(cen[order], now, 1.0, 10.0)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts[5:]) * 1e3
print(f"mot_tracks_step, {T} centroids / {out['n_tracks']} tracks, L = 40: median {np.median(ts):.3f} ms  min {ts.min():.3f} ms  max {ts.max():.3f} ms  launches {trk.last_launches()}")
trk.set_profiling(True)
for f in range(frames, frames + 5):
    now = 0.1 * f
    cen = np.zeros((T, 4), np.float32)
    cen[:, :2] = base + vel * now
    cen[:, 3] = now
    trk.tracks_step(cen[rng.permutation(T)], now, 1.0, 10.0)
for k, (ms, c) in sorted(trk.profile().items(), key=lambda kv: -kv[1][0]):
    print(f"    {k:24s} {ms / c * 1e3:9.1f} us x{c}")
