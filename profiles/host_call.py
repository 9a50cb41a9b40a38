"""Wall clock of the one-frame HOST entry (pageable numpy buffers in and out; mot_frame = removeStatic + clustering + tables +
circumcentres) on a c1 frame, with and without the pinned staging:  python profiles/host_call.py [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as entry

mot = entry.load_package()
synth = mot.synth
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 60
occ, res, origin = synth.make_map_c1()
cloud, _ = synth.make_frame_c1()
p = synth.C1_PARAMS
for stage in ("1", "0", "1", "0"):
    os.environ["MOT_HOST_STAGE"] = stage
    trk = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
    trk.set_map(occ, res, origin[:2], static_tolarance=p["static_tolerance"])
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    bufs = trk.frame_buffers(len(cloud))
    for _ in range(5):
        M, K = trk.frame_into(cloud, bufs, 1.0)
    ts = []
    for _ in range(calls):
        t0 = time.perf_counter()
        M, K = trk.frame_into(cloud, bufs, 1.0)
        ts.append(time.perf_counter() - t0)
    print(f"MOT_HOST_STAGE={stage}  mot_frame, c1 ({len(cloud)} points in, {M} kept, {K} clusters out), caller-owned buffers: median {np.median(ts) * 1e6:7.1f} us  min {np.min(ts) * 1e6:7.1f} us  launches {trk.last_launches()}")
    trk.close()
