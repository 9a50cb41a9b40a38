"""A few c1 frames through the small-frame path (for ncu captures): python profiles/small_frame_once.py [calls] [full|cluster]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

mot = entry.load_package()
synth = mot.synth
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
full = (sys.argv[2] if len(sys.argv) > 2 else "full") == "full"
occ, res, origin = synth.make_map_c1()
cloud, _ = synth.make_frame_c1()
p = synth.C1_PARAMS
trk = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
trk.set_map(occ, res, origin[:2], static_tolarance=p["static_tolerance"])
trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
d = torch.from_numpy(cloud).cuda()
for _ in range(calls):
    trk.frame_device(d.data_ptr(), len(cloud), full, full, 1.0)
print(trk.result_counts(), trk.small_frames(), {k: round(v / 1e3, 1) for k, v in trk.small_frame_phases().items()})
