"""Multi-GPU plumbing for frame batches: frames are independent units, so rank r owns a contiguous
frame range and the only exchange is a gather of the per-frame cluster tables to rank 0.

torch.distributed is used strictly as plumbing (NCCL on GPUs, gloo in the CPU tests); nothing here touches the
data path of a frame.  Imported lazily by bench.py / tests -- the product library itself has no torch dependency.
"""
import numpy as np


def frame_range(rank, world_size, n_frames):
    """Contiguous frame range [lo, hi) owned by `rank` (GPU g gets frames [g*F/G, (g+1)*F/G))."""
    lo = (rank * n_frames) // world_size
    hi = ((rank + 1) * n_frames) // world_size
    return lo, hi


def pack_tables(tables):
    """tables: list (one per local frame) of K_f x 10 float32 arrays (count, mean, bbox_min, bbox_max).
    Returns (counts int64[F_local], payload float32[sum K_f, 10])."""
    counts = np.array([len(t) for t in tables], dtype=np.int64)
    payload = np.concatenate([np.asarray(t, dtype=np.float32).reshape(-1, 10) for t in tables]) if len(tables) and counts.sum() else np.zeros((0, 10), np.float32)
    return counts, payload


def gather_tables(counts, payload, device=None, group=None):
    """Gathers every rank's (counts, payload) on rank 0.  counts: int64 tensor [F_local] (same F_local on every
    rank), payload: float32 tensor [rows, 10] on `device`.  Returns on rank 0 a list over ranks of
    (counts, payload[:rows_r]) tensors; None elsewhere.  Two collectives: all_gather of the row counts (so every
    rank can size the padded payload), then gather of the padded payload."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = device if device is not None else payload.device
    counts = counts.to(device)
    rows = torch.tensor([payload.shape[0]], dtype=torch.int64, device=device)
    all_rows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(all_rows, rows, group=group)
    max_rows = int(max(int(r.item()) for r in all_rows))
    padded = torch.zeros((max(max_rows, 1), 10), dtype=torch.float32, device=device)
    if payload.shape[0]:
        padded[: payload.shape[0]] = payload
    all_counts = [torch.zeros_like(counts) for _ in range(world)] if rank == 0 else None
    all_payload = [torch.zeros_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(counts, all_counts, dst=0, group=group)
    dist.gather(padded, all_payload, dst=0, group=group)
    if rank != 0:
        return None
    return [(all_counts[r], all_payload[r][: int(all_rows[r].item())]) for r in range(world)]


def gather_table_block(block, gather_list=None, group=None):
    """One collective for a whole block of steps: `block` is a float32 tensor [G, rows + 1, 10] -- per step, row 0 holds the
    number of valid table rows in column 0, the table follows -- of the same shape on every rank.  Rank 0 passes
    gather_list (world tensors shaped like block) and finds every rank's block there afterwards.  Used by bench.py so
    that the exchange costs one NCCL launch per G steps instead of three per step (a collective's kernel has to wait
    for SM space behind the frame kernels, so few large gathers beat many small ones)."""
    import torch.distributed as dist

    dist.gather(block, gather_list if dist.get_rank(group) == 0 else None, dst=0, group=group)
    return gather_list


def unpack_table_block(block):
    """Inverse of the packing above for one rank's block: list over steps of [K, 10] tensors."""
    out = []
    for j in range(block.shape[0]):
        k = int(block[j, 0, 0].item())
        out.append(block[j, 1:k + 1])
    return out
