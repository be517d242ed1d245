"""C5-shaped run through the per-shard entry (lrge_gpu_twoset_counts_shard) under torchrun: every rank generates ONLY
its block of targets (README.md:173-175's human read set is 2 000 000 targets: 250 000 per rank on 8 ranks) and all
queries, nobody ever holds the whole target set.  No oracle can run at this size; the check is the estimate itself --
the median per-read estimate must come out at the simulated genome size.

  torchrun --nproc-per-node W tests/manual/c5_shard_run.py [targets_per_rank] [n_queries] [mean_len] [genome_size]
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lrge_b200  # noqa: E402
from lrge_b200 import readsim  # noqa: E402


def main():
    a = sys.argv[1:]
    per_rank = int(a[0]) if len(a) > 0 else 250_000
    nq = int(a[1]) if len(a) > 1 else 25_000
    mean_len = int(a[2]) if len(a) > 2 else 12_000
    gsize = int(a[3]) if len(a) > 3 else 3_000_000_000
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl")
    t0 = time.time()
    g = readsim.genome(gsize)
    blk = readsim.reads(g, per_rank, mean_len, readsim.ONT, readsim.SEED_TARGET + 7919 * rank, f"t{rank}_")
    Q = readsim.reads(g, nq, mean_len, readsim.ONT, readsim.SEED_QUERY, "q")
    del g
    t_gen = time.time() - t0
    ctx = lrge_b200.Context(dev)
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(lrge_b200.Context.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(ident, 0)
    ctx.comm_init(rank, world, bytes(ident.cpu().numpy().tobytes()))
    tot = torch.tensor([blk.n, blk.total_bases], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    n_targets, target_bases = int(tot[0]), int(tot[1])
    runs = []
    for it in range(2):                                   # the first call allocates
        dist.barrier(); torch.cuda.synchronize()
        t1 = time.time()
        res = ctx.twoset_counts_shard(blk, Q, "ont", shard_index=rank, shard_count=world)
        torch.cuda.synchronize(); dist.barrier()
        runs.append(time.time() - t1)
    free, total = torch.cuda.mem_get_info()
    if rank == 0:
        lens = Q.lengths()
        avg_t = np.float32(target_bases) / np.float32(n_targets)
        est = np.array([lrge_b200.per_read_estimate(int(lens[i]), float(avg_t), n_targets, int(res.counts[i]), 100) for i in range(Q.n)],
                       dtype=np.float32)
        lo, md, up = lrge_b200.estimate_from(est, True, 0.15, 0.65)
        st = res.stats
        print(json.dumps({"workload": f"C5-shaped: {world} ranks x {per_rank} targets ({target_bases / 1e9:.2f} Gbases), {nq} queries, "
                                      f"{mean_len} b mean, genome {gsize}",
                          "seconds_per_call": runs, "host_generation_s": round(t_gen, 1),
                          "estimate": md, "lower": lo, "upper": up, "estimate_over_truth": (md or 0) / gsize,
                          "overlaps": int(res.counts.sum()), "queries_without_overlap": int((res.counts == 0).sum()),
                          "mid_occ": st["mid_occ"], "n_keys": st["n_keys"], "n_target_minimizers_all_ranks": st["n_target_minimizers"],
                          "n_anchors_rank0": st["n_anchors"], "stage_ms_rank0": {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")},
                          "hbm_used_gb_rank0": round((total - free) / 1e9, 1)}), flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
