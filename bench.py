#!/usr/bin/env python
"""bench.py -- Euclidean-clustering throughput of the B200 hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # one process per GPU under torchrun for N > 1
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload = "c2"): BASELINE config[1] -- 2^20-point 128-beam-style synthetic frames, cluster
tolerance 0.5 m, min 5 / max 100000.  A "step" is one pass of the hot path (voxel grid build, union-find,
size filter + CSR emission, per-cluster table) over a batch of `--frames` DISTINCT frames per GPU, one after the
other; with the default 8 frames the step's inputs are 128 MiB > the 126 MB L2, so no frame is served from a
warm L2.  N > 1: every rank clusters its own frames (weak scaling) and rank 0 gathers the cluster tables.

value  = points/s with the frames already resident in HBM (mot_cluster_batch_device; CUDA events on the handle's stream);
         the median of --reps timed regions of exactly --steps steps each (all regions are listed in `timed_regions_ms`).
e2e    = the same metric through the reference-facing call mot_cluster_batch() with pinned HOST buffers: H2D copy of
         the cloud and D2H copy of the CSR result inside the timed region; `h2d_ceiling` is the pinned-copy bandwidth of
         all ranks copying at once on this box, `frac_of_ceiling` says how much of it the call reaches.
configs = the other BASELINE configs (c1, c3, c4, c5) at full size on rank 0, each with its parity against the oracle
         checked in the same run (N = 1 only; --no-configs skips it).
The CPU oracle (oracle/) is only used for the parity checks, the cpu_baseline leg and for --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

# what one step of `--impl reference` clusters (a full step of the GPU arm would take the CPU path > 10 minutes per step)
REFERENCE_SAMPLE = ("--impl reference clusters one 1/16 azimuth wedge (~65k points) of a c2 frame per host thread and step; "
                    "once per run it also times one full 2^20-point frame on one thread (reported as full_frame)")
METRIC = "euclidean_clustering_throughput"
UNIT = "Mpoints/s"


def make_config(p, n_pts, F, world):
    """Identical for both arms (the driver compares them)."""
    return {"workload": "c2", "points_per_frame": n_pts, "frames_per_step_per_gpu": F, "cluster_tolerance": p["cluster_tolerance"],
            "min_cluster_size": p["min_cluster_size"], "max_cluster_size": p["max_cluster_size"],
            "l2_policy": f"{F} distinct frames per step ({F * n_pts * 16 >> 20} MiB of input > 126 MB L2)",
            "parallelism": f"frames sharded over {world} GPU(s), tables gathered to rank 0" if world > 1 else "single GPU",
            "reference_sample": REFERENCE_SAMPLE}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms from the start of the timed region to the end of the GPU work."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.f.close()
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
            sm = [float(r[0]) for r in rows if len(r) >= 8]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][1])
                out["samples"] = len(sm)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nme in enumerate(names):
                    if any(r[4 + i].strip() == "Active" for r in rows if len(r) >= 8):
                        out["reasons"].append(nme)
        except Exception:
            pass
        return out


def kernel_bytes(name, M, cf, cc, K, total, key_bytes, n_sort_passes, R_G=1024 * 592):
    """ALGORITHMIC bytes one launch of `name` must move (per-point figures follow SURVEY 8d)."""
    kb = key_bytes
    table = {
        "k_bbox": 16 * M,
        "k_cell_keys": (16 + kb) * M,
        "k_rs_hist[cells]": kb * M,                       # onesweep: one read of the keys for every pass's histogram
        "k_rs_scatter[cells]": 2 * (kb + 4) * M,          # one onesweep pass: read pair + write pair
        "k_cells_count": kb * M,
        "k_cells_write": (kb + 4 + 16 + 16) * M + 24 * cf + 16 * cc,  # key, index, gather point, write sorted SoA point, tables
        "k_coarse_records": 12 * cf + 80 * cc,            # record (16 B) + neighbour row (64 B) per coarse cell
        "k_uf_sparse": 16 * M + 80 * cc + 8 * cf,         # every sorted point once + record/neighbour row + parent r/w
        "k_uf_sparse2": 16 * M + 80 * cc + 8 * cf,        # same traffic, half a warp per coarse cell
        "k_hash_build": 4 * cc + 16 * cc,                    # coarse key in, one hash slot (key + value, <= 50 % load) out
        "k_cell_local": 16 * M + 41 * cf + 56 * cc,         # every sorted point once; cell tables in, boxes / records / parents out
        "k_uf_fused": 36 * cc + 40 * cf,                    # keys, records, hash, boxes, parents (points only for ambiguous pairs)
        "k_uf_cross": 68 * cc + 40 * cf,
        "k_uf_survivors": 36 * cc,
        "k_uf_walk": 40 * cf,
        "k_compact_onepass<map>": 32 * M,                   # overwritten by the caller with 16 N + 16 M where N is known
        "k_uf_flatten<in-place>": 8 * cf,
        "k_uf_flatten<root>": 8 * cf,
        "k_comp_accumulate": 16 * cf,
        "k_kept_list": 12 * cf + 12 * K,
        "k_clusters_small": 24 * K,
        "k_point_rank": (16 + 4 + 4 + 4) * M,             # sorted point (cell id in .w), index, key out, label out
        "k_rs_hist[csr]": 4 * M,
        "k_rs_scatter[csr]": 16 * M,
        "k_stats_accumulate": 20 * total + 48 * K,
        # v1 kernels (MOT_UF_MODE=0)
        "k_uf_pairs<1>": 16 * M + 16 * cf + 12 * cc,
        "k_uf_pairs<2>": 16 * cf + 12 * cc,
    }
    return table.get(name)


def run_b200(args, rank, world, local_rank):
    import torch

    mot = entry.load_package()
    synth = mot.synth
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    from importlib import import_module  # shard helpers live in the package
    shard = import_module("mot_b200.shard")

    p = synth.C2_PARAMS
    scene = synth.scene_c2()
    F = args.frames
    frames_np = [scene.frame(rank * F + f) for f in range(F)]
    n_pts = len(frames_np[0])
    all_np = np.ascontiguousarray(np.concatenate(frames_np))
    frame_offsets = np.arange(F + 1, dtype=np.int64) * n_pts
    d_all = torch.from_numpy(all_np).to(dev)          # the step's frames, resident in HBM
    # handles per GPU: 4 saturate the GPU (2: -8 %, 3: -2.4 %, 6: +0.1 %, 8: -1 %; profiles/r02_streams.txt); 3 when the ranks of this node
    # would otherwise run more host threads (handles + the gather thread) than there are cores
    S = args.streams if args.streams > 0 else (4 if world * 5 <= (os.cpu_count() or 32) else 3)
    # every handle has a host thread that waits on its stream.  Spinning (the default) is fastest while each thread has a
    # core; when the ranks of this node together run more threads than cores, the waits poll-and-yield instead
    # (MOT_SYNC=yield, read by mot_create): 37.2 vs 34.4 Gpoints/s at 8 GPUs on 32 cores against spinning with 3 streams
    if "MOT_SYNC" not in os.environ and world * (S + 1) > (os.cpu_count() or 32):
        os.environ["MOT_SYNC"] = "yield"
    trks = [mot.Tracker(device=local_rank, max_points=F * n_pts, max_tracks=0) for _ in range(S)]
    for t_ in trks:
        t_.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    trk = trks[0]
    TABLE_ROWS = 1 << 15
    n_slots = 2 * S
    slots = [torch.zeros((TABLE_ROWS, 10), dtype=torch.float32, device=dev) for _ in range(n_slots)]
    slot_rows = [0] * n_slots
    slot_full = [threading.Event() for _ in range(n_slots)]
    slot_free = [threading.Event() for _ in range(n_slots)]
    # the steps' tables travel to rank 0 in blocks of GATHER_BLOCK steps: one collective per block (shard.gather_table_block)
    GATHER_BLOCK = 8
    agg = torch.zeros((GATHER_BLOCK, TABLE_ROWS + 1, 10), dtype=torch.float32, device=dev) if world > 1 else None
    agg_all = [torch.zeros_like(agg) for _ in range(world)] if world > 1 and rank == 0 else None

    def step_device(gather, s=0, step_index=0):
        """One pass of the hot path over the step's batch of F frames (frame id rides in the voxel key).  With several
        GPUs the step's cluster table is left in a slot for the gather (done by the main thread in step order)."""
        tk = trks[s]
        tk.cluster_batch_device(d_all.data_ptr(), frame_offsets)
        # the step's per-cluster table (count, mean, bbox) is part of the path at every N: written to a device slot
        sl = step_index % n_slots
        if gather:
            slot_free[sl].wait()
            slot_free[sl].clear()
        M, K, total = tk.result_counts()
        k = min(K, TABLE_ROWS)
        if k:
            tk.lib.mot_result_fetch(tk.h, None, 0, None, 0, None, 0, slots[sl].data_ptr(), None, k)
        launches = tk.last_launches()  # the handle counts every kernel launched since the step began
        slot_rows[sl] = k
        if gather:
            slot_full[sl].set()
        return launches

    def run_steps(fn, n_steps, gather=False):
        """n_steps steps spread round-robin over the S handles (one host thread + one CUDA stream each): while one
        batch waits on a host round trip or a PCIe copy, the others keep the SMs busy.  With gather=True the main
        thread gathers every step's table to rank 0 in step order (same collective order on every rank)."""
        for e in slot_free:
            e.set()
        for e in slot_full:
            e.clear()
        acc = [0] * S

        def worker(s):
            for i in range(s, n_steps, S):
                acc[s] += fn(s, i)

        th = [threading.Thread(target=worker, args=(s,)) for s in range(S)]
        for t_ in th:
            t_.start()
        if gather:
            for i in range(n_steps):
                sl = i % n_slots
                slot_full[sl].wait()
                slot_full[sl].clear()
                k = slot_rows[sl]
                j = i % GATHER_BLOCK
                agg[j, 0, 0] = float(k)
                if k:
                    agg[j, 1:k + 1].copy_(slots[sl][:k], non_blocking=True)
                torch.cuda.current_stream().synchronize()  # the slot may be refilled once its rows are in the block
                slot_free[sl].set()
                if j == GATHER_BLOCK - 1 or i == n_steps - 1:
                    shard.gather_table_block(agg, agg_all)
        for t_ in th:
            t_.join()
        return sum(acc)

    gather = world > 1
    run_steps(lambda s, i: step_device(gather, s, i), max(args.warmup, 3) * S, gather)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall0 = time.perf_counter()
    regions = []
    launches = 0
    for rep in range(max(1, args.reps)):  # every region: exactly --steps steps between a barrier + synchronize on both sides
        trk.timer_start()
        n_l = run_steps(lambda s, i: step_device(gather, s, i), args.steps, gather)
        ms = trk.timer_stop()
        torch.cuda.synchronize()
        if rep == 0:
            launches = n_l
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        regions.append(float(t[0]))
    wall_ms = (time.perf_counter() - t_wall0) * 1e3 / max(1, args.reps)
    t = torch.tensor([wall_ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, wall_max = float(np.median(regions)), float(t[0])
    pts_per_step = world * F * n_pts
    value = pts_per_step * args.steps / (ms_max * 1e-3) / 1e6

    # ---- e2e: host buffers through mot_cluster_batch (pinned), copies inside the timed region ----
    import ctypes as C
    h_all = torch.from_numpy(all_np).pin_memory()
    h_out = [(torch.empty(F + 1, dtype=torch.int32).pin_memory(), torch.empty(F * n_pts + 1, dtype=torch.int32).pin_memory(),
              torch.empty(F * n_pts, dtype=torch.int32).pin_memory()) for _ in range(S)]

    def step_e2e(s=0, step_index=0):
        h_fco, h_off, h_idx = h_out[s]
        kk = C.c_int32(0)
        tk = trks[s]
        rc = tk.lib.mot_cluster_batch(tk.h, h_all.data_ptr(), frame_offsets, F, h_fco.data_ptr(), h_off.data_ptr(), F * n_pts + 1,
                                      h_idx.data_ptr(), F * n_pts, C.byref(kk))
        assert rc == 0, tk.lib.mot_last_error(tk.h)
        return 4 * (F + 1) + 4 * (kk.value + 1) + 4 * int(h_off[kk.value])

    run_steps(step_e2e, 3 * S)
    if dist:
        dist.barrier()
    e2e_steps = max(2 * S, min(args.steps, 12))
    trk.timer_start()
    d2h_bytes = run_steps(step_e2e, e2e_steps)
    e_ms = trk.timer_stop()
    te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
    if dist:
        dist.barrier()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = pts_per_step * e2e_steps / (float(te[0]) * 1e-3) / 1e6

    # ---- the same step with PACKED 12-byte points (x, y, z without pcl::PointXYZ's padding word; mot_frame_batch expands them on the
    # device): a quarter fewer PCIe bytes.  Reported beside `e2e`, not as it: the reference's in-memory cloud is the 16-byte layout.
    h_all12 = torch.from_numpy(np.ascontiguousarray(all_np[:, :3])).pin_memory()

    def step_e2e12(s=0, step_index=0):
        h_fco, h_off, h_idx = h_out[s]
        kk = C.c_int32(0)
        tk = trks[s]
        rc = tk.lib.mot_frame_batch(tk.h, h_all12.data_ptr(), 12, frame_offsets, F, 0, None, None, h_fco.data_ptr(), h_off.data_ptr(), F * n_pts + 1,
                                    h_idx.data_ptr(), F * n_pts, C.byref(kk), None, None, 0)
        assert rc == 0, tk.lib.mot_last_error(tk.h)
        return 0

    run_steps(step_e2e12, 2 * S)
    if dist:
        dist.barrier()
    trk.timer_start()
    run_steps(step_e2e12, e2e_steps)
    e12_ms = trk.timer_stop()
    te12 = torch.tensor([e12_ms], dtype=torch.float64, device=dev)
    if dist:
        dist.barrier()
        dist.all_reduce(te12, op=dist.ReduceOp.MAX)
    e2e12_value = pts_per_step * e2e_steps / (float(te12[0]) * 1e-3) / 1e6
    del h_all12

    # ---- what the box can copy: every rank copies its pinned step input to its GPU at the same time (same bytes as a step) ----
    d_sink = torch.empty_like(d_all)
    for _ in range(2):
        d_sink.copy_(h_all, non_blocking=True)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(4):
        d_sink.copy_(h_all, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    tc = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if dist:
        dist.barrier()
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    h2d_gbs_per_gpu = 4 * F * n_pts * 16 / (float(tc[0]) * 1e-3) / 1e9
    h2d_ceiling = world * h2d_gbs_per_gpu * 1e9 / 16 / 1e6  # Mpoints/s if the step were nothing but its input copy
    del d_sink

    # ---- the same frames through the per-frame call the reference's callback would make (mot_cluster, one frame per
    # call, pinned host buffers, H2D + D2H inside the timed region), frames dealt round-robin to the S handles ----
    small = [mot.Tracker(device=local_rank, max_points=n_pts, max_tracks=0) for _ in range(S)]
    for t_ in small:
        t_.set_cluster_params(p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"])
    h_frames = [h_all[f * n_pts:(f + 1) * n_pts] for f in range(F)]

    def frame_e2e(s=0, i=0):
        _, h_off, h_idx = h_out[s]
        kk = C.c_int32(0)
        tk = small[s]
        rc = tk.lib.mot_cluster(tk.h, h_frames[i % F].data_ptr(), n_pts, h_off.data_ptr(), n_pts + 1, h_idx.data_ptr(), n_pts, C.byref(kk))
        assert rc == 0, tk.lib.mot_last_error(tk.h)
        return 0

    run_steps(frame_e2e, 2 * F)
    if dist:
        dist.barrier()
    n_calls = 6 * F
    trk.timer_start()
    run_steps(frame_e2e, n_calls)
    pf_ms = trk.timer_stop()
    tpf = torch.tensor([pf_ms], dtype=torch.float64, device=dev)
    if dist:
        dist.barrier()
        dist.all_reduce(tpf, op=dist.ReduceOp.MAX)
    e2e_per_frame = world * n_calls * n_pts / (float(tpf[0]) * 1e-3) / 1e6
    for t_ in small:
        t_.close()

    # ---- single-frame latency (one 2^20-point frame per call, as the reference's callback sees it) ----
    lat = []
    for f in range(F):
        trk.frame_device(d_all[f * n_pts:(f + 1) * n_pts].data_ptr(), n_pts)
    for rep in range(3):
        for f in range(F):
            trk.timer_start()
            trk.frame_device(d_all[f * n_pts:(f + 1) * n_pts].data_ptr(), n_pts)
            lat.append(trk.timer_stop())
    single_frame_us = float(np.median(lat)) * 1e3
    clocks = sampler.stop()

    # ---- per-kernel attribution (separate pass with event pairs around every launch) ----
    trk.set_profiling(True)
    prof_steps = 3
    for _ in range(prof_steps):
        step_device(False, 0, 0)
    prof = trk.profile()
    trk.set_profiling(False)
    F_prof = 1  # kernel figures are per launch over the whole batch
    M, K, total = trk.result_counts()
    grid = trk.result_grid()
    key_bytes = 4 if grid["key_bits"] <= 32 else 8
    peak, peak_src = peaks()
    kernels = []
    tot_kernel_ms = sum(v[0] for v in prof.values())
    for name, (tms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        avg_ms = tms / cnt
        b = kernel_bytes(name, M, grid["fine_cells"], grid["coarse_cells"], K, total, key_bytes, 0)
        kernels.append({"kernel": name, "launches_per_step": cnt / prof_steps, "avg_us": round(avg_ms * 1e3, 2),
                        "share": round(tms / tot_kernel_ms, 4), "alg_bytes": b,
                        "gbs": round(b / (avg_ms * 1e-3) / 1e9, 1) if b else None})
    top = kernels[0]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if tj.get("frames_per_launch") == args.frames:  # the capture is per launch: only valid for the same batch size
            traffic = tj.get(top["kernel"])
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                "frac": round(top["gbs"] / peak, 4) if top["gbs"] else None, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": top["alg_bytes"], "avg_launch_us": top["avg_us"], "share_of_kernel_time": top["share"]}
    # whole-frame figure against SURVEY 8d's B_frame (no removeStatic at c2: N-term dropped)
    P = (grid["key_bits"] + 9) // 10
    b_frame = (152 + 16 * P) * M + 16 * grid["coarse_cells"] + 44 * K
    step_us = ms_max * 1e3 / args.steps
    whole = {"alg_bytes_per_step": b_frame, "step_us": round(step_us, 1), "gbs": round(b_frame / (step_us * 1e-6) / 1e9, 1),
             "frac": round(b_frame / (step_us * 1e-6) / 1e9 / peak, 4)}

    # north_star's sub-target: grid build + union-find against (104 + 16 P) M + 16 C bytes
    k15 = ("k_cell_keys", "k_rs_hist[cells]", "k_rs_scan[cells]", "k_rs_scatter[cells]", "k_cells_count", "k_hash_clear", "k_cells_write",
           "k_coarse_records", "k_uf_sparse", "k_uf_sparse2", "k_uf_dense<1>", "k_uf_dense<2>", "k_uf_flatten<in-place>", "k_uf_flatten<root>",
           "k_uf_pairs<1>", "k_uf_pairs<2>", "k_cell_local", "k_uf_fused", "k_uf_cross", "k_uf_survivors", "k_uf_walk", "k_uf_heavy<1>",
           "k_uf_heavy<2>", "k_uf_flatten_if", "k_cell_local_dense", "k_hash_build")
    t15_us = sum(v[0] for kname, v in prof.items() if kname in k15) / prof_steps * 1e3
    b15 = (104 + 16 * P) * M + 16 * grid["coarse_cells"]
    grid_uf = {"alg_bytes_per_step": b15, "kernel_us_per_step": round(t15_us, 1), "gbs": round(b15 / (t15_us * 1e-6) / 1e9, 1),
               "frac": round(b15 / (t15_us * 1e-6) / 1e9 / peak, 4), "target_frac": 0.8}

    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_max / args.steps, 4), "timed_regions_ms": [round(r, 3) for r in regions], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": make_config(p, n_pts, F, world),
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": F * n_pts * 16, "d2h_bytes_per_step": d2h_bytes // e2e_steps,
                "api": "mot_cluster_batch (host pinned buffers in, CSR out)",
                "h2d_ceiling": {"value": round(h2d_ceiling, 1), "unit": UNIT, "gbs_per_gpu": round(h2d_gbs_per_gpu, 2),
                                "how": "all ranks copy their pinned step input to their GPU at once (cudaMemcpyAsync, CUDA events, max over ranks)"},
                "frac_of_ceiling": round(e2e_value / h2d_ceiling, 3),
                "per_frame_api": {"value": round(e2e_per_frame, 2), "unit": UNIT, "api": "mot_cluster, one frame per call"},
                "packed_xyz12": {"value": round(e2e12_value, 2), "unit": UNIT, "h2d_bytes_per_step": F * n_pts * 12,
                                 "api": "mot_frame_batch, point_stride_bytes = 12 (same frames without the padding word of pcl::PointXYZ)"}},
        "single_frame_latency_us": round(single_frame_us, 1),
        "streams_per_gpu": S, "host_wait": os.environ.get("MOT_SYNC", "spin"),
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "frame_roofline": whole,
        "grid_build_union_find_roofline": grid_uf,
        "kernels": kernels[:16],
        "wall_ms_per_step": round(wall_max / args.steps, 4),
        "result": {"kept_points": M, "clusters": K, "indices": total, **grid},
    }
    if rank == 0 and world == 1 and not args.no_configs:
        import bench_configs
        out["configs"] = bench_configs.run(mot, entry.load_oracle(), local_rank, peak, quick=args.quick_configs)
        c1 = out["configs"].get("c1", {})
        # the reference's own frame shape (c1: 65,536 points + map) beside the 2^20-point figure above: CUDA events around the call, and
        # first-to-last kernel instruction on the device
        out["small_frame_latency_us"] = {"around_call": round(c1["ms"] * 1e3, 1) if "ms" in c1 else None, "on_device": c1.get("device_us"),
                                         "host_buffers": round(c1["host_buffers_ms"] * 1e3, 1) if "host_buffers_ms" in c1 else None,
                                         "workload": "c1 frame, removeStatic + clustering + tables + circumcentres"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(frames_np[0], p)
    for t_ in trks:
        t_.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else None


def wedge(frame, frac):
    az = np.arctan2(frame[:, 1], frame[:, 0])
    return np.ascontiguousarray(frame[az < -np.pi + 2 * np.pi * frac])


def cpu_baseline(frame, p):
    """The oracle's restatement of the reference path (KD-tree built twice + BFS, 1 thread -- the reference is single
    threaded, MOT.cpp:117-121) on a 162-degree azimuth wedge of one frame of the step (bounded: ~15-20 s of CPU work; the
    cost is far from linear in the wedge -- the densely sampled near-range surfaces dominate it -- so the figure is
    quoted together with the sample it was taken on)."""
    oracle = entry.load_oracle()
    w = wedge(frame, 0.45)
    t0 = time.perf_counter()
    off, idx = oracle.cluster_kdtree(w, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"], build_twice=True)
    dt = time.perf_counter() - t0
    return {"value": round(len(w) / dt / 1e6, 4), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"162-degree azimuth wedge of frame 0 ({len(w)} points, {len(off) - 1} clusters), {dt:.1f} s, oracle KD-tree+BFS restatement of PCL"}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path.  PCL/FLANN cannot be built here, so this is the oracle
    port (kind "port"): KD-tree (leaf 15) built twice + sorted radius search + serial BFS, one frame wedge per host
    thread (frames are independent), all host threads."""
    if rank != 0:
        return None
    from concurrent.futures import ThreadPoolExecutor

    mot = entry.load_package()
    oracle = entry.load_oracle()
    synth = mot.synth
    p = synth.C2_PARAMS
    scene = synth.scene_c2()
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64) - 1)  # one core is left to the full-frame timing that runs beside the steps
    base = [wedge(scene.frame(f), 1.0 / 16) for f in range(min(threads, 8))]
    work = [base[i % len(base)] for i in range(threads)]
    pts = sum(len(w) for w in work)

    def one(w):
        return oracle.cluster_kdtree(w, p["cluster_tolerance"], p["min_cluster_size"], p["max_cluster_size"], build_twice=True)

    # the step's real unit, once: a whole 2^20-point frame on one thread (the reference's callback is single threaded,
    # MOT.cpp:117-121).  The per-core cost grows faster than the point count, so the wedge figure flatters the CPU.
    full = {}

    def full_frame():
        fr = scene.frame(0)
        t0 = time.perf_counter()
        off, _ = one(fr)
        dt = time.perf_counter() - t0
        full.update({"points": len(fr), "seconds": round(dt, 2), "clusters": len(off) - 1, "value_per_core": round(len(fr) / dt / 1e6, 5), "unit": UNIT})

    th_full = threading.Thread(target=full_frame)
    if not args.no_full_frame:
        th_full.start()

    def step():
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, work))

    warmup = max(args.warmup, 3)  # same rule as the GPU arm
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    if not args.no_full_frame:
        th_full.join()
    value = pts * args.steps / dt / 1e6
    sample = f"{threads} x 1/16 azimuth wedges of c2 frames ({pts} points per step), one wedge per host thread; oracle KD-tree + BFS restatement of PCL"
    return {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(p, 1 << 20, args.frames, world),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "full_frame": full or None},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=16, help="distinct frames per step per GPU")
    ap.add_argument("--streams", type=int, default=0, help="handles (host thread + CUDA stream each) per GPU working on alternate steps; 0 = choose")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-frame", action="store_true", help="reference arm: skip the one-off full-frame timing (~90 s)")
    ap.add_argument("--no-configs", action="store_true", help="skip the c1 / c3 / c4 / c5 block (N = 1 only)")
    ap.add_argument("--quick-configs", action="store_true", help="smaller c3 / c4 shapes in the configs block (smoke runs)")
    ap.add_argument("--reps", type=int, default=3, help="timed regions of --steps steps each; value is their median")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        out = run_reference(args, rank, world)
    else:
        out = run_b200(args, rank, world, local_rank)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
