"""Single-call latency (host wall clock around the C-ABI call, device-resident input, median of N calls) with the speculative
grid plan on and off:  python profiles/latency.py [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as entry

mot = entry.load_package()
synth = mot.synth
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 60


def wall_us(fn):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(calls):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e6, float(np.min(ts)) * 1e6


cases = []
p2 = synth.C2_PARAMS
cases.append(("c2 frame, 1,048,576 points, clustering only", synth.scene_c2().frame(0), p2, False))
p1 = synth.C1_PARAMS
c1, _ = synth.make_frame_c1()
cases.append(("c1 frame, 65,536 points, clustering only", c1, p1, False))
cases.append(("c1 frame, 65,536 points, removeStatic + clustering + centroids", c1, p1, True))
occ, res, origin = synth.make_map_c1()
for name, cloud, p, full in cases:
    d = torch.from_numpy(cloud).cuda()
    trk = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    if full:
        trk.set_map(occ, res, origin[:2], static_tolarance=p1["static_tolerance"])
    for spec in (True, False, True, False):
        trk.grid_plan(spec)
        med, best = wall_us(lambda: trk.frame_device(d.data_ptr(), len(cloud), full, full, 1.0))
        print(f"{name:66s} plan_spec={int(spec)}  median {med:7.1f} us  min {best:7.1f} us  launches {trk.last_launches()}  hits/misses {trk.grid_plan()}")
    if trk.last_launches() <= 7:
        print("    small-frame phases (us):", {k: round(v / 1e3, 1) for k, v in trk.small_frame_phases().items()}, trk.small_frames(), trk.result_counts(), trk.result_grid())
    trk.close()
