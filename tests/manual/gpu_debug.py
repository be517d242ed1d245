"""Stage-by-stage diagnostics on a GPU box (not a test): prints where the CUDA path first departs from the oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lrge_b200  # noqa: E402
import oracle  # noqa: E402
from lrge_b200 import readsim  # noqa: E402


def diff_minimizers(ctx, rs, platform, tag):
    x, y, cnt = ctx.dump_minimizers(rs, platform)
    want, wcnt = oracle.sketch_readset(rs, oracle.options(platform))
    ok = x.shape[0] == want.shape[0] and np.array_equal(x, want["x"]) and np.array_equal(y, want["y"])
    print(f"[{tag}] minimizers gpu={x.shape[0]} oracle={want.shape[0]} equal={ok}")
    if not ok:
        bad = np.nonzero(cnt != wcnt)[0]
        print("  reads with different counts:", bad[:10], "of", bad.shape[0])
        n = min(x.shape[0], want.shape[0])
        d = np.nonzero((x[:n] != want["x"][:n]) | (y[:n] != want["y"][:n]))[0]
        if d.shape[0]:
            i = int(d[0])
            print("  first diff at", i, "gpu", hex(int(x[i])), hex(int(y[i])), "oracle", hex(int(want['x'][i])), hex(int(want['y'][i])))
            for j in range(max(0, i - 2), min(n, i + 4)):
                print("   ", j, "gpu rid/pos/z", int(y[j]) >> 32, (int(y[j]) & 0xffffffff) >> 1, int(y[j]) & 1, hex(int(x[j])),
                      "| orc", int(want['y'][j]) >> 32, (int(want['y'][j]) & 0xffffffff) >> 1, int(want['y'][j]) & 1, hex(int(want['x'][j])))
    return ok


def main():
    ctx = lrge_b200.Context(0)
    for platform, err in (("ont", readsim.ONT), ("pb", readsim.HIFI)):
        g = readsim.genome(300_000)
        T = readsim.reads(g, 300, 4000, err, readsim.SEED_TARGET, "t")
        Q = readsim.reads(g, 120, 4000, err, readsim.SEED_QUERY, "q")
        diff_minimizers(ctx, T, platform, platform)
        opt = oracle.options(platform)
        idx = oracle.Index(T, opt, threads=8)
        mid, nk = ctx.dump_mid_occ(T, platform)
        print(f"[{platform}] mid_occ gpu={mid} oracle={opt.mid_occ} keys gpu={nk} oracle={idx.n_keys}")
        t0 = time.time()
        res = ctx.twoset_counts(T, Q, platform)
        t1 = time.time()
        want, had, st = oracle.twoset_counts(T, Q, platform, threads=8)
        print(f"[{platform}] twoset gpu {t1 - t0:.3f}s stats", res.stats)
        print(f"[{platform}] oracle stats", st)
        bad = np.nonzero(res.counts != want)[0]
        print(f"[{platform}] counts equal={bad.shape[0] == 0} mismatches={bad.shape[0]} had_equal={np.array_equal(res.had_mapping, had)}")
        for q in bad[:5]:
            print("   q", q, "gpu", res.counts[q], "oracle", want[q])
    for name in sys.argv[1:]:
        strat, platform, A, B = readsim.make_config(name)
        t0 = time.time()
        if strat == "twoset":
            res = ctx.twoset_counts(A, B, platform)
            t1 = time.time()
            want, had, st = oracle.twoset_counts(A, B, platform, threads=os.cpu_count())
        else:
            res = ctx.ava_counts(A, platform)
            t1 = time.time()
            want, st = oracle.ava_counts(A, platform, threads=os.cpu_count())
        t2 = time.time()
        bad = np.nonzero(res.counts != want)[0]
        print(f"[{name}] gpu {t1 - t0:.3f}s oracle {t2 - t1:.3f}s counts equal={bad.shape[0] == 0} mismatches={bad.shape[0]}")
        print(f"[{name}] gpu stats", res.stats)
        print(f"[{name}] oracle stats", st)
    ctx.close()


if __name__ == "__main__":
    main()
