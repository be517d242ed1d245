"""Run under torchrun with 2+ GPUs: the sharded index (full target set on every rank) and the per-shard entry (each rank
holds only its block) must both return the oracle's counts.  Used by tests/test_multi_gpu.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lrge_b200  # noqa: E402
from lrge_b200 import readsim  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    g = readsim.genome(300_000)
    T = readsim.reads(g, 900, 4000, readsim.ONT, readsim.SEED_TARGET, "t")
    Q = readsim.reads(g, 300, 4000, readsim.ONT, readsim.SEED_QUERY, "q")
    ctx = lrge_b200.Context(int(os.environ.get("LOCAL_RANK", rank)))
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(lrge_b200.Context.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(ident, 0)
    ctx.comm_init(rank, world, bytes(ident.cpu().numpy().tobytes()))
    full = ctx.twoset_counts(T, Q, "ont", shard_index=rank, shard_count=world)
    # unequal blocks on purpose: the entry accepts any partition
    cuts = [0] + [int(T.n * (0.3 + 0.7 * (r + 1) / world)) if r + 1 < world else T.n for r in range(world)]
    cuts[1] = max(cuts[1], 1)
    blk = T.subset(np.arange(cuts[rank], cuts[rank + 1]))
    part = ctx.twoset_counts_shard(blk, Q, "ont", shard_index=rank, shard_count=world)
    ok = True
    if rank == 0:
        sys.path.insert(0, ROOT)
        import oracle
        rc, rh, rst = oracle.twoset_counts(T, Q, "ont", threads=8)
        for name, got in (("sharded", full), ("per-shard entry", part)):
            same = bool(np.array_equal(got.counts, rc) and np.array_equal(got.had_mapping.astype(bool), rh.astype(bool)))
            print(f"{name}: counts identical = {same}, mid_occ {got.stats['mid_occ']} vs {rst['mid_occ']}, overlaps {int(got.counts.sum())}", flush=True)
            ok = ok and same and got.stats["mid_occ"] == rst["mid_occ"]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
