#!/usr/bin/env python
"""bench.py -- read-pair overlaps/sec of the overlap hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (liblrge_b200.so through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A step is one pass of the whole hot path (sketch -> index -> seed lookup -> chaining -> per-read counts)
over the two-set workload BASELINE.json's metric is quoted on (configs[1]: 5 000 query x 10 000 target
synthetic ONT reads, mean 8 kb, 4.4 Mbp genome).  With N > 1 (one process per GPU under torchrun) the
target set grows to N x 10 000 reads and is sharded in N contiguous blocks across ranks -- per-GPU work stays fixed
("weak") -- and every step ends with the library's NCCL all-reduce of the per-read counts.

Prints ONE JSON line on rank 0.  `value` is timed with the read sets already resident in HBM; `e2e` goes
through the host-buffer C-ABI call with pinned inputs, H2D and D2H inside the timed region.  Keys the driver keeps:
  roofline      the dominant kernel against the measured HBM peak, algorithmic bytes per SURVEY.md 8(d); also carries
                `kernels` (live per-kernel CUDA-event table), `stage_ms` and `counts` of the last step
  cpu_baseline  the oracle port on this box's host cores; also carries `parity` (counts identical to the GPU's over
                ALL reads) and `configs`: the other BASELINE.json configurations (C1, C3, C4) run and verified here
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "read_pair_overlaps_per_sec"
UNIT = "overlaps/s"
WORKLOAD = "C2"


def workload(name: str, n_gpus: int, scale: float):
    from lrge_b200 import readsim
    strat, platform, T, Q = readsim.make_config(name, scale=scale, target_mult=n_gpus)
    assert strat == "twoset"
    desc = (f"{name} two-set {Q.n} query x {T.n} target synthetic ONT reads, mean 8 kb, "
            f"{readsim.CONFIGS[name][4] * scale / 1e6:.2f} Mbp genome")
    return platform, T, Q, desc


def config_dict(desc: str):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": desc, "targets_scale_with_gpus": True, "sharding": "N contiguous target blocks, one per GPU",
            "l2": "256 MiB memset between timed steps; index + hit lists (> 0.5 GB) exceed L2"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md), sampled through NVML so that no
    nvidia-smi process is spawned next to the measurement; falls back to nvidia-smi if NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.samples.append((float(sm), float(mx), [k for k, b in bits.items() if r & b]))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.samples.append((float(f[0]), float(f[1]), [names[i] for i in range(4) if f[3 + i].lower().startswith("active")]))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.1 if self.nvml else 0.5)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        sm = [s[0] for s in self.samples]
        mx = [s[1] for s in self.samples]
        reasons = sorted({r for s in self.samples for r in s[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "via": "nvml" if self.nvml else "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args):
    """The reference's own CPU path for this workload: the oracle port (minimap2 2.30 restatement + liblrge
    counting; the Rust/minimap2-sys build is not possible in this image), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    platform, T, Q, desc = workload(args.config, args.gpus, args.scale)
    threads = os.cpu_count() or 1
    # Bounded sample: at N = 1 the whole workload; at N > 1 (targets x N) the full target set is indexed but only the
    # first Q/N queries are mapped, and the whole-workload time is EXTRAPOLATED as t_index + t_map * Q/Qs (index build
    # does not depend on the queries; mapping cost is per query).
    qs = Q.n if args.gpus <= 1 else max(250, Q.n // args.gpus)
    Qs = Q if qs == Q.n else Q.subset(range(qs))
    scale_q = Q.n / qs
    times, ovl = [], 0.0
    for i in range(args.warmup + args.steps):
        counts, had, st = oracle.twoset_counts(T, Qs, platform, threads=threads)
        dt = st["t_index_s"] + st["t_map_s"] * scale_q
        if i >= args.warmup:
            times.append(dt)
        ovl = float(counts.sum()) * scale_q
    ms = 1e3 * sum(times) / len(times)
    value = ovl / (ms / 1e3)
    sample = ("the full workload (index build + map + count), every step; measured" if qs == Q.n else
              f"EXTRAPOLATED: all {T.n} targets indexed, first {qs} of {Q.n} queries mapped; step time = t_index + t_map * {scale_q:.1f}, "
              f"overlaps scaled by {scale_q:.1f} (quote the N = 1 line for measured-vs-measured ratios)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": config_dict(desc),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "extrapolated": qs != Q.n, "overlaps_per_step": ovl,
                         "gbases_per_s": (T.total_bases + Q.total_bases) / 1e9 / (ms / 1e3)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def other_configs(ctx, names, threads, rank, world, dist, shard):
    """BASELINE.json's other configurations on this box: GPU time (CUDA events, 2 passes after a warm-up) and parity of
    the per-read counts over ALL reads against the oracle on the host cores."""
    import lrge_b200
    import oracle
    from lrge_b200 import readsim
    out = {}
    for name in names:
        t0 = time.perf_counter()
        strat, platform, A, B = readsim.make_config(name)
        gen_s = time.perf_counter() - t0
        # read sets resident in HBM, like `value`; time = CUDA events around the whole path (2 passes after a warm-up)
        sA = ctx.stage(A)
        sB = ctx.stage(B) if B is not None else None
        run = (lambda: ctx.twoset_counts(sA, sB, platform, **shard)) if strat == "twoset" else (lambda: ctx.ava_counts(sA, platform, **shard))
        res = run()
        ms = []
        for _ in range(2):
            res = run()
            ms.append(res.stats["ms_total"])
        sA.free()
        if sB is not None:
            sB.free()
        e = {"strategy": strat, "platform": platform, "n_gpus": world, "reads": A.n + (B.n if B is not None else 0),
             "mbases": (A.total_bases + (B.total_bases if B is not None else 0)) / 1e6,
             "gpu_ms": round(float(np.mean(ms)), 3), "generate_s": round(gen_s, 1),
             "n_anchors": res.stats["n_anchors"], "n_overlaps": res.stats["n_overlaps"], "rounds": res.stats["n_rounds"], "batches": res.stats["n_batches"],
             "deferred_queries": res.stats["n_deferred"], "mid_occ": res.stats["mid_occ"],
             "stage_ms": {k[3:]: round(v, 2) for k, v in res.stats.items() if k.startswith("ms_") and v}}
        if world > 1:
            t = __import__("torch").tensor([e["gpu_ms"]], dtype=__import__("torch").float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e["gpu_ms"] = float(t[0])
        e["overlaps_per_s"] = e["n_overlaps"] / (e["gpu_ms"] / 1e3)
        if rank == 0:
            t0 = time.perf_counter()
            if strat == "twoset":
                want, had, st = oracle.twoset_counts(A, B, platform, threads=threads)
                same = bool(np.array_equal(want, res.counts) and np.array_equal(had, res.had_mapping))
            else:
                want, st = oracle.ava_counts(A, platform, threads=threads)
                same = bool(np.array_equal(want, res.counts))
            e.update({"oracle_s": round(time.perf_counter() - t0, 1), "oracle_threads": threads,
                      "counts_identical": same, "reads_compared": int(want.shape[0]),
                      "anchors_identical": bool(st["n_anchors"] == res.stats["n_anchors"]) if world == 1 else None,
                      "speedup_vs_oracle": round((time.perf_counter() - t0) * 1e3 / e["gpu_ms"], 1)})
        if dist is not None:
            dist.barrier()
        out[name] = e
        del A, B, res
    return out


def run_ours(args):
    import torch
    import lrge_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: liblrge_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=30))

    platform, T, Q, desc = workload(args.config, world, args.scale)
    ctx = lrge_b200.Context(local)
    if world > 1:
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident.copy_(torch.frombuffer(bytearray(lrge_b200.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, 0)
        ctx.comm_init(rank, world, bytes(ident.cpu().numpy().tobytes()))
    shard = dict(shard_index=rank, shard_count=world)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sT, sQ = ctx.stage(T), ctx.stage(Q)

    # ---- value: inputs resident in HBM ---------------------------------------------------------------
    last = None
    for _ in range(args.warmup):
        last = ctx.twoset_counts(sT, sQ, platform, **shard)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev_ms, launches, step_ms = 0.0, 0, []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        last = ctx.twoset_counts(sT, sQ, platform, **shard)
        ev_ms += last.stats["ms_total"]
        step_ms.append(round(last.stats["ms_total"], 3))
        launches += last.stats["n_kernel_launches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.summary() if rank == 0 else None
    t = torch.tensor([ev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ev_ms, wall_ms = float(t[0]), float(t[1])
    ms_per_step = ev_ms / args.steps
    overlaps = int(last.counts.sum())          # counts are already reduced over all shards
    value = overlaps / (ms_per_step / 1e3)

    # ---- per-kernel CUDA-event pass (EVERY rank: the sharded path contains NCCL collectives) ------------------
    ctx.set_profiling(True)
    prof_steps = 3
    acc = {}
    for i in range(prof_steps + 1):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.twoset_counts(sT, sQ, platform, **shard)
        if i == 0:
            continue                                   # first pass creates the events
        for kt in ctx.kernel_times():
            a = acc.setdefault(kt["name"], dict(ms=0.0, bytes=0.0, launches=0, passes=kt.get("passes", 1)))
            a["ms"] += kt["ms"]; a["bytes"] += kt["algorithmic_bytes"]; a["launches"] += kt["launches"]
    ctx.set_profiling(False)
    # ---- e2e: host buffers through the C-ABI, H2D + D2H inside ------------------------------------------
    pT, bT = lrge_b200.pinned_copy(T)
    pQ, bQ = lrge_b200.pinned_copy(Q)
    for _ in range(min(args.warmup, 2)):
        ctx.twoset_counts(pT, pQ, platform, **shard)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_last = ctx.twoset_counts(pT, pQ, platform, **shard)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    assert np.array_equal(e2e_last.counts, last.counts)
    # what the link gives a plain pinned copy of the same size (context for e2e - value): one cudaMemcpyAsync, CUDA events
    h2d_gbs = None
    try:
        nb = int(T.off[T.n * (rank + 1) // world] - T.off[T.n * rank // world]) + int(Q.total_bases)
        host = torch.empty(nb, dtype=torch.uint8).pin_memory()
        devb = torch.empty(nb, dtype=torch.uint8, device="cuda")
        devb.copy_(host, non_blocking=True); torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); devb.copy_(host, non_blocking=True); e1.record(); torch.cuda.synchronize()
        h2d_gbs = nb / 1e9 / (e0.elapsed_time(e1) / 1e3)
        del host, devb
    except Exception:
        pass
    lo, hi = T.n * rank // world, T.n * (rank + 1) // world          # a rank copies its own target block + all queries
    h2d_rank = int(T.off[hi] - T.off[lo]) + int(Q.total_bases)
    h2d_all = int(T.total_bases) + world * int(Q.total_bases)
    d2h = int(5 * Q.n)

    threads = os.cpu_count() or 1
    # ---- parity + CPU baseline on rank 0: the oracle port on this box's cores -------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        import oracle
        t0 = time.perf_counter()
        oc, oh, ost = oracle.twoset_counts(T, Q, platform, threads=threads)
        dt = time.perf_counter() - t0
        cpu = {"value": int(oc.sum()) / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"the full N = {world} workload once (index build + map + count)", "seconds": dt,
               "parity": {"counts_identical_to_gpu": bool(np.array_equal(oc, last.counts)),
                          "had_mapping_identical": bool(np.array_equal(oh, last.had_mapping)),
                          "reads_compared": int(oc.shape[0]), "overlaps": int(oc.sum())}}
    if dist is not None:
        dist.barrier()
    # ---- the other BASELINE.json configurations, verified over all reads (N = 1: C1, C3, C4; N = 4: C4) -------
    names = [n for n in args.configs.split(",") if n and n != "none"] if args.configs != "auto" else \
        (["C1", "C3", "C4"] if world == 1 else ["C4"] if world == 4 else [])
    if args.no_cpu_baseline:
        names = []
    sT.free()
    sQ.free()
    cfgs = other_configs(ctx, names, threads, rank, world, dist, shard) if names else {}

    if rank != 0:
        ctx.close()
        dist.destroy_process_group()
        return
    st = last.stats
    # ---- roofline of the dominant kernel (rank 0): algorithmic bytes per SURVEY.md 8(d) ------------------------------
    peak, peak_src = measured_peak()
    kernels = []
    for name, a in acc.items():
        ms = a["ms"] / prof_steps
        gbs = a["bytes"] / 1e9 / (a["ms"] / 1e3) if a["ms"] > 0 else 0.0
        k = {"name": name, "ms_per_step": round(ms, 4), "launches_per_step": a["launches"] / prof_steps,
             "algorithmic_bytes_per_step": a["bytes"] / prof_steps, "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        if a["passes"] > 1:                      # a library radix sort: what one of its passes moves per second
            k["passes"] = a["passes"]
            k["per_pass_gbs"] = round(gbs * a["passes"], 1)
        kernels.append(k)
    kernels.sort(key=lambda k: -k["ms_per_step"])
    dom = kernels[0]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("workload") == args.config:
                traffic = tj.get("dram_bytes_per_launch", {}).get(dom["name"])
        except Exception:
            pass
    ksum = sum(k["ms_per_step"] for k in kernels)
    roofline = {"kernel": dom["name"], "bound": "hbm",
                "achieved": dom["algorithmic_bytes_per_step"] / dom["launches_per_step"] / 1e9 /
                            (dom["ms_per_step"] / dom["launches_per_step"] / 1e3),
                "peak": peak, "unit": "GB/s", "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_step"] / dom["launches_per_step"],
                "kernel_ms_per_launch": dom["ms_per_step"] / dom["launches_per_step"],
                "share_of_step": dom["ms_per_step"] / ms_per_step,
                "whole_step": {"algorithmic_bytes": sum(k["algorithmic_bytes_per_step"] for k in kernels),
                               "gbs": sum(k["algorithmic_bytes_per_step"] for k in kernels) / 1e9 / (ms_per_step / 1e3),
                               "kernel_ms_sum": round(ksum, 4), "kernel_time_share_of_step": round(ksum / ms_per_step, 4)},
                "note": "algorithmic bytes per SURVEY.md 8(d): a sort counts one read + one write of its records; the fused "
                        "map kernel counts the K4 hit-list + anchor bytes, K5 and K6 (8H + 64A + 4NQ) although its anchors "
                        "never leave shared memory, so `traffic` (ncu DRAM bytes) is far below it",
                "kernels": kernels,
                "stage_ms": {k: st[k] for k in st if k.startswith("ms_")},
                "counts": {k: st[k] for k in ("n_target_minimizers", "n_query_minimizers", "n_keys", "n_seeds", "n_hits", "n_anchors",
                                              "n_pairs", "mid_occ", "n_batches", "n_rounds", "n_deferred")}}
    roofline["frac"] = roofline["achieved"] / peak
    roofline["whole_step"]["frac"] = roofline["whole_step"]["gbs"] / peak
    if cpu is not None:
        cpu["configs"] = cfgs
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": config_dict(desc),
        "overlaps_per_step": overlaps,
        "gbases_per_s": (T.total_bases + Q.total_bases) / 1e9 / (ms_per_step / 1e3),
        "candidate_pairs_per_s": T.n * Q.n / (ms_per_step / 1e3),
        "wall_ms_per_step": wall_ms / args.steps, "step_ms": step_ms,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": overlaps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h * world,
                "h2d_bytes_per_step_per_rank": h2d_rank, "d2h_bytes_per_step_per_rank": d2h, "ms_per_step": e2e_ms,
                "plain_pinned_h2d_gbs_rank0": h2d_gbs},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    for b in (bT, bQ):
        b.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=WORKLOAD)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every oracle run (profiling passes)")
    ap.add_argument("--configs", default="auto", help="other BASELINE.json configurations to run and verify "
                    "(comma list; auto = C1,C3,C4 at N=1, C4 at N=4, none otherwise; '' = none)")
    ap.add_argument("--verify", action="store_true", help="kept for compatibility: parity is always checked now")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    need = [os.path.join(ROOT, "lrge_b200", "csrc", "liblrge_b200.so"), os.path.join(ROOT, "lrge_b200", "csrc", "libreadsim.so"),
            os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__ as g
            g.build()
        else:
            while not all(os.path.exists(p) for p in need):
                time.sleep(1.0)
            time.sleep(2.0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
