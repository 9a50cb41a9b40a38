"""Generates tests/golden/sim_01_occupancy.npz from the reference-owned map fixture.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_sim01_fixture.py
The .npz holds the nav_msgs/OccupancyGrid that ROS map_server would publish for map/sim_01.{pgm,yaml}
(trinary rule) -- i.e. exactly what ObstacleTrack::mapCallback (MOT.cpp:235-251) receives.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = "/root/reference/map"
occ, res, origin = oracle.load_map_server_trinary(os.path.join(REF, "sim_01.pgm"), os.path.join(REF, "sim_01.yaml"))
vals, counts = np.unique(occ, return_counts=True)
print("shape", occ.shape, "res", res, "origin", origin, dict(zip(vals.tolist(), counts.tolist())))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sim_01_occupancy.npz"), occ=occ, resolution=np.float32(res),
                    origin=np.array(origin, dtype=np.float64))
