"""CPU tests of the multi-GPU plumbing: contiguous frame ranges and the gather of per-frame cluster
tables to rank 0, exercised with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest


def test_frame_ranges_partition(mot):
    import importlib
    shard = importlib.import_module("mot_b200.shard")
    for n_frames in (1, 7, 64, 512):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.frame_range(r, world, n_frames)
                assert 0 <= lo <= hi <= n_frames
                seen += list(range(lo, hi))
            assert seen == list(range(n_frames))
    assert shard.frame_range(3, 8, 512) == (192, 256)  # 64 frames per GPU


def _worker(rank, world, port, q):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as entry
    entry.load_package()
    shard = importlib.import_module("mot_b200.shard")
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    tables = [rng.random((k, 10), dtype=np.float32) for k in ((3, 0, 5) if rank == 0 else (2, 7, 1))]
    counts, payload = shard.pack_tables(tables)
    res = shard.gather_tables(torch.from_numpy(counts), torch.from_numpy(payload), device=torch.device("cpu"))
    # block-wise exchange used by bench.py: G steps per collective, row 0 of every step carries its row count
    G, rows = 3, 8
    block = torch.zeros((G, rows + 1, 10))
    for j, t in enumerate(tables):
        block[j, 0, 0] = float(len(t))
        block[j, 1:len(t) + 1] = torch.from_numpy(t)
    gl = [torch.zeros_like(block) for _ in range(world)] if rank == 0 else None
    shard.gather_table_block(block, gl)
    if rank == 0:
        out = [(c.numpy().tolist(), p.numpy()) for c, p in res]
        blocks = [[t.numpy() for t in shard.unpack_table_block(b)] for b in gl]
        q.put((out, payload, blocks))
    else:
        assert res is None
        q.put(("payload", rank, payload))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_tables_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rank0 = [g for g in got if g[0] != "payload"][0]
    rank1 = [g for g in got if g[0] == "payload"][0]
    out, payload0, blocks = rank0
    for r, pay in ((0, payload0), (1, rank1[2])):   # the block gather delivers the same tables, step by step
        assert [len(t) for t in blocks[r]] == out[r][0] and np.array_equal(np.concatenate(blocks[r]), pay)
    assert out[0][0] == [3, 0, 5] and out[1][0] == [2, 7, 1]
    assert np.array_equal(out[0][1], payload0) and np.array_equal(out[1][1], rank1[2])
    assert out[1][1].shape == (10, 10)
