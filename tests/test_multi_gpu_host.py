"""world_size-2 `gloo` test of the multi-GPU scheme on CPU: contiguous target blocks [n*r/W, n*(r+1)/W) -- the product's
shard_range -- are additive under GLOBAL occurrence filtering, and the reduction the library does with NCCL (sum of
counts, max of had_mapping) reproduces the single-process result.  There is no GPU here, so the oracle stands in for the
per-shard compute (tests only); the product's own sharded path is checked on hardware by bench.py at every N > 1
(cpu_baseline.parity) and by tests/test_gpu_parity.py::test_target_sharding_is_additive."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from lrge_b200 import readsim

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = readsim.genome(150_000)
    T = readsim.reads(g, 160, 3000, readsim.ONT, readsim.SEED_TARGET, "t")
    Q = readsim.reads(g, 60, 3000, readsim.ONT, readsim.SEED_QUERY, "q")
    counts, had, st = oracle.twoset_counts(T, Q, "ont", threads=2, shard_index=rank, shard_count=world)
    c = torch.from_numpy(counts.astype(np.int64))
    h = torch.from_numpy(had.astype(np.int64))
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(h, op=dist.ReduceOp.MAX)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # max-over-ranks timing reduce used by bench.py
    if rank == 0:
        full, full_had, full_st = oracle.twoset_counts(T, Q, "ont", threads=2)
        # the block boundaries of the product (lrge_gpu.cu shard_range): every target belongs to exactly one rank
        blocks = [(T.n * r // world, T.n * (r + 1) // world) for r in range(world)]
        out["ok_blocks"] = blocks[0][0] == 0 and blocks[-1][1] == T.n and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        out["ok_counts"] = bool(np.array_equal(c.numpy(), full.astype(np.int64)))
        out["ok_had"] = bool(np.array_equal(h.numpy(), full_had.astype(np.int64)))
        out["ok_mid"] = st["mid_occ"] == full_st["mid_occ"]
        out["ok_max"] = float(t) == float(world)
        out["overlaps"] = int(full.sum())
    dist.destroy_process_group()


def test_two_rank_sharded_counts_reduce_to_full():
    world = 2
    port = 29000 + os.getpid() % 2000
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res["ok_counts"] and res["ok_had"] and res["ok_mid"] and res["ok_max"] and res["ok_blocks"]
    assert res["overlaps"] > 0
