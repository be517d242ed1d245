"""External confirmation of the (unpinned) oracle against REAL minimap2 -- run wherever a minimap2 2.30 binary exists
(none in the build image).

    python tests/manual/confirm_with_minimap2.py [--minimap2 /path/to/minimap2] [--config C1] [--scale 0.2] [--platform ont]

Writes target.fa / query.fa (what liblrge hands to minimap2, twoset.rs:157-186), runs
    minimap2 -x ava-ont --dual=yes -I 9999G -t N target.fa query.fa        (two-set; liblrge: dual=true, batch_size=max)
    minimap2 -x ava-ont            -I 9999G -t N reads.fa  reads.fa         (all-vs-all)
and compares, per query read, the set of distinct target names (PAF column 6) -- liblrge's HashSet<target_name>
(twoset.rs:286-302) -- and every PAF row (columns 1-11 + cm + s1) with the oracle's regs.
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile
from collections import defaultdict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from lrge_b200 import readsim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minimap2", default=shutil.which("minimap2"))
    ap.add_argument("--config", default="C1")
    ap.add_argument("--scale", type=float, default=0.2)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    if not a.minimap2:
        raise SystemExit("no minimap2 binary found: pass --minimap2 (needs v2.30 to match minimap2-sys 0.1.30+minimap2.2.30)")
    print(subprocess.run([a.minimap2, "--version"], capture_output=True, text=True).stdout.strip(), "(expected 2.30)")
    strat, platform, A, B = readsim.make_config(a.config, scale=a.scale)
    preset = "ava-ont" if platform == "ont" else "ava-pb"
    tmp = tempfile.mkdtemp(prefix="lrge_confirm_")
    if strat == "twoset":
        A.write_fasta(os.path.join(tmp, "target.fa"))
        B.write_fasta(os.path.join(tmp, "query.fa"))
        cmd = [a.minimap2, "-x", preset, "--dual=yes", "-I", "9999G", "-t", str(a.threads), os.path.join(tmp, "target.fa"), os.path.join(tmp, "query.fa")]
        T, Q, dual = A, B, True
    else:
        A.write_fasta(os.path.join(tmp, "reads.fa"))
        cmd = [a.minimap2, "-x", preset, "-I", "9999G", "-t", str(a.threads), os.path.join(tmp, "reads.fa"), os.path.join(tmp, "reads.fa")]
        T, Q, dual = A, A, False
    paf = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    got_sets, got_rows = defaultdict(set), set()
    for line in paf.splitlines():
        f = line.split("\t")
        tags = {t[:2]: t[5:] for t in f[12:]}
        got_sets[f[0]].add(f[5])
        got_rows.add((f[0], int(f[2]), int(f[3]), f[4], f[5], int(f[7]), int(f[8]), int(f[9]), int(f[10]), int(tags["cm"]), int(tags["s1"])))
    regs, qidx, mid_occ = oracle.twoset_regs(T, Q, platform, threads=a.threads, dual=dual)
    want_sets, want_rows = defaultdict(set), set()
    for r, q in zip(regs, qidx):
        qn, tn = Q.name(int(q)).decode(), T.name(int(r["rid"])).decode()
        want_sets[qn].add(tn)
        want_rows.add((qn, int(r["qs"]), int(r["qe"]), "-" if r["rev"] else "+", tn, int(r["rs"]), int(r["re"]), int(r["mlen"]), int(r["blen"]),
                       int(r["cnt"]), int(r["score"])))
    names = [Q.name(i).decode() for i in range(Q.n)]
    bad = [n for n in names if got_sets.get(n, set()) != want_sets.get(n, set())]
    print(f"oracle mid_occ {mid_occ}; reads whose distinct-target sets differ: {len(bad)} of {len(names)}")
    print(f"PAF rows: minimap2 {len(got_rows)}, oracle {len(want_rows)}, only minimap2 {len(got_rows - want_rows)}, only oracle {len(want_rows - got_rows)}")
    for n in bad[:5]:
        print("  ", n, "minimap2-only", sorted(got_sets[n] - want_sets[n])[:5], "oracle-only", sorted(want_sets[n] - got_sets[n])[:5])
    print("work dir:", tmp)
    sys.exit(1 if bad or got_rows != want_rows else 0)


if __name__ == "__main__":
    main()
