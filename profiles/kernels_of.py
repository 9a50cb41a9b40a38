"""Per-kernel table of one call shape: python profiles/kernels_of.py c1|c2frame|c4:<tol>  (A/B of builds: MOT_B200_LIB=...)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as entry

mot = entry.load_package()
synth = mot.synth
what = sys.argv[1] if len(sys.argv) > 1 else "c1"
reps = 10
if what == "c1":
    occ, res, origin = synth.make_map_c1()
    cloud, _ = synth.make_frame_c1()
    p = synth.C1_PARAMS
    trk = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
    trk.set_map(occ, res, origin[:2], static_tolarance=p["static_tolerance"])
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    d = torch.from_numpy(cloud).cuda()
    call = lambda: trk.frame_device(d.data_ptr(), len(cloud), True, True, 1.0)
elif what == "c2frame":
    p = synth.C2_PARAMS
    cloud = synth.scene_c2().frame(0)
    trk = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    d = torch.from_numpy(cloud).cuda()
    call = lambda: trk.frame_device(d.data_ptr(), len(cloud))
else:
    tol = float(what.split(":")[1])
    p = synth.C4_PARAMS
    cloud = synth.make_frame_c4()
    trk = mot.Tracker(device=0, max_points=len(cloud), max_tracks=0)
    trk.set_cluster_params(tol, p["min_cluster_size"], p["max_cluster_size"])
    d = torch.from_numpy(cloud).cuda()
    call = lambda: trk.frame_device(d.data_ptr(), len(cloud))
for _ in range(3):
    call()
ts = []
for _ in range(reps):
    trk.timer_start()
    call()
    ts.append(trk.timer_stop())
trk.set_profiling(True)
for _ in range(reps):
    call()
prof = trk.profile()
tot = sum(ms for ms, c in prof.values()) / reps
cnt = trk.result_counters()
print(f"=== {what} [{os.environ.get('MOT_ENV_NOTE', '')}]: call {np.median(ts) * 1e3:.1f} us, sum of kernels {tot * 1e3:.1f} us, launches {trk.last_launches()}, counts {trk.result_counts()} "
      f"grid {trk.result_grid()} heavy {cnt[9]}/{cnt[10]} timings {trk.timings()}")
for k, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOPK", "14"))]:
    print(f"    {k:28s} {ms / c * 1e3:9.1f} us x{c / reps:g}")
