"""A/B of union-find / grid-build variants on one c2 batch: per-kernel table, internal counters, and the component labels
compared with a second variant on the same GPU (bit-exact or the script fails).

    python profiles/exp_uf.py [frames] [reps] -- "MOT_UF_MODE=2" "MOT_UF_MODE=2 MOT_UF_LIGHT=16" ...

The first variant listed is the label baseline.  Environment switches are read by mot_create.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as entry

mot = entry.load_package()
synth = mot.synth
args = sys.argv[1:]
variants = ["MOT_UF_MODE=1"]
if "--" in args:
    i = args.index("--")
    variants = args[i + 1:]
    args = args[:i]
F = int(args[0]) if len(args) > 0 else 16
reps = int(args[1]) if len(args) > 1 else 5
workload = os.environ.get("EXP_WORKLOAD", "c2")
if workload == "c2":
    p = synth.C2_PARAMS
    sc = synth.scene_c2()
    frames = [sc.frame(f) for f in range(F)]
elif workload == "c3":
    p = synth.C3_PARAMS
    sc = synth.scene_c3()
    frames = [sc.frame(f) for f in range(F)]
else:
    raise SystemExit("EXP_WORKLOAD = c2 | c3")
n = len(frames[0])
allp = np.ascontiguousarray(np.concatenate(frames))
fo = np.zeros(F + 1, dtype=np.int64)
fo[1:] = np.cumsum([len(f) for f in frames])
d = torch.from_numpy(allp).cuda()
base_labels = None
TOPK = int(os.environ.get("TOPK", "12"))
for var in variants:
    env = dict(kv.split("=", 1) for kv in var.split())
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        trk = mot.Tracker(device=0, max_points=len(allp), max_tracks=0)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    trk.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    for _ in range(3):
        trk.cluster_batch_device(d.data_ptr(), fo)
    # unprofiled wall time of the call (stream timer)
    ts = []
    for _ in range(reps):
        trk.timer_start()
        trk.cluster_batch_device(d.data_ptr(), fo)
        ts.append(trk.timer_stop())
    trk.debug_stats()
    trk.cluster_batch_device(d.data_ptr(), fo)
    st = trk.debug_stats()
    if st:
        print("    stats of one call:", st)
    trk.set_profiling(True)
    for _ in range(reps):
        trk.cluster_batch_device(d.data_ptr(), fo)
    prof = trk.profile()
    trk.set_profiling(False)
    tot = sum(ms for ms, c in prof.values()) / reps
    cnt = trk.result_counters()
    print(f"=== {var}: counts {trk.result_counts()} grid {trk.result_grid()} launches {trk.last_launches()}")
    print(f"    call {np.median(ts) * 1e3:.1f} us (min {min(ts) * 1e3:.1f}); sum of kernels {tot * 1e3:.1f} us; heavy1 {cnt[9]} heavy2 {cnt[10]} serial-fallback {cnt[12]} flags {cnt[4]}")
    for k, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])[:TOPK]:
        print(f"    {k:28s} {ms / c * 1e3:10.1f} us x{c / reps:g}")
    lab = trk.result_labels()
    if base_labels is None:
        base_labels = lab
    else:
        same = np.array_equal(lab, base_labels)
        print(f"    labels identical to '{variants[0]}': {same}")
        if not same:
            bad = np.flatnonzero(lab != base_labels)
            print(f"    MISMATCH at {len(bad)} points, first {bad[:10]}")
            sys.exit(1)
    trk.close()
print("exp ok")
