"""Runs a few c2 frames through mot_frame_device (for ncu captures): python profiles/one_frame.py [n_frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

mot = entry.load_package()
synth = mot.synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
p = synth.C2_PARAMS
fr = synth.scene_c2().frame(0)
d = torch.from_numpy(fr).cuda()
trk = mot.Tracker(device=0, max_points=len(fr), max_tracks=0)
trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
for _ in range(2):  # warm-up (lazy module loading, allocations)
    trk.frame_device(d.data_ptr(), len(fr))
trk.set_profiling(True)
for _ in range(n):
    trk.frame_device(d.data_ptr(), len(fr))
print(trk.result_counts(), trk.result_grid())
for k, (ms, c) in sorted(trk.profile().items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} {ms / c * 1e3:10.1f} us x{c}")
