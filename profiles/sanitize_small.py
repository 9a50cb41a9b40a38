"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): KAT clouds, a c1 frame, a batch, IHGP."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import __graft_entry__ as entry
from cases import kat_cases

os.environ.setdefault("MOT_UF_MODE", "2")  # the round-2 cell path (these clouds are far below the size where it is the default)
mot = entry.load_package()
synth = mot.synth
trk = mot.Tracker(device=0, max_points=1 << 17, max_tracks=64)
for name, (pts, tol, mn, mx, _) in sorted(kat_cases().items()):
    trk.set_cluster_params(tol, mn, mx)
    off, idx = trk.extract(pts)
    if len(off) > 1:
        trk.cluster_stats()
        trk.get_centroid(1.0)
occ, res, origin = synth.make_map_c1()
cloud, _ = synth.make_frame_c1(n_points=32768)
trk.set_map(occ, res, origin[:2], static_tolarance=2)
trk.set_cluster_params(0.3, 5, 300)
out = trk.frame(cloud, 1.0)
trk.voxel_grid(cloud, (0.1, 0.1, 2.0))
sc = synth.scene_c3()
trk.extract_batch([sc.frame(f, n_points=20000) for f in range(3)])
dense = synth.make_frame_c4(n_points=1 << 15, n_blobs=12)
trk.set_cluster_params(1.0, 5, 100000)
trk.extract(dense)
trk.set_cluster_params(0.3, 5, 300)
clouds = [synth.make_frame_c1(n_points=9000 + 100 * f, frame=f)[0] for f in range(3)]
trk.frame_batch(clouds, do_remove_static=True, stamps=[1.0, 1.1, 1.2], packed12=True)
raw = np.ascontiguousarray(cloud[:, :3]).view(np.uint8).reshape(len(cloud), 12)
trk.cluster_pointcloud2(raw, len(cloud), 12, (0, 4, 8), voxel_leaf_size=0.05, do_remove_static=True, stamp_minus_time_init=2.0)
rings = synth.make_rings_c5(48, 12)
hyp = (np.exp(-5.5), np.exp(-3.5), np.exp(0.75))
trk.ihgp_configure(0.1, 0.03, hyp, hyp, 12)
trk.ihgp_step(rings, np.zeros((48, 4)))
trk.close()
print("sanitize run ok", out["K"])
