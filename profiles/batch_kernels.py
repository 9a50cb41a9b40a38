"""Per-kernel table of one c2 batch through mot_cluster_batch_device (A/B of library builds: MOT_B200_LIB=...):
python profiles/batch_kernels.py [frames] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as entry

mot = entry.load_package()
synth = mot.synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
p = synth.C2_PARAMS
sc = synth.scene_c2()
frames = [sc.frame(f) for f in range(F)]
n = len(frames[0])
allp = np.ascontiguousarray(np.concatenate(frames))
fo = np.arange(F + 1, dtype=np.int64) * n
d = torch.from_numpy(allp).cuda()
trk = mot.Tracker(device=0, max_points=len(allp), max_tracks=0)
trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
for _ in range(3):
    trk.cluster_batch_device(d.data_ptr(), fo)
trk.set_profiling(True)
for _ in range(reps):
    trk.cluster_batch_device(d.data_ptr(), fo)
prof = trk.profile()
tot = sum(ms for ms, c in prof.values()) / reps
print(trk.result_counts(), f"kernel time per batch {tot * 1e3:.1f} us")
for k, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOPK", "6"))]:
    print(f"{k:28s} {ms / c * 1e3:10.1f} us x{c}")
