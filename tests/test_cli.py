"""lrge-b200, the CLI-compatible driver (lrge/src/cli.rs flags) over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

import lrge_b200
from conftest import ROOT, has_gpu
from lrge_b200 import readsim

BIN = os.path.join(ROOT, "lrge_b200", "csrc", "lrge-b200")


def test_cli_help_and_bad_args():
    r = subprocess.run([BIN, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--filter-contained" in r.stdout and "--use-min-ref" in r.stdout
    assert subprocess.run([BIN], capture_output=True).returncode == 2
    assert subprocess.run([BIN, "--q1", "0.9", "x.fa"], capture_output=True).returncode == 2   # cli.rs validate_low_quantile
    # cli.rs: --num conflicts with --target / --query; the long quiet / verbose flags exist
    r = subprocess.run([BIN, "-n", "5", "-T", "3", "x.fa"], capture_output=True, text=True)
    assert r.returncode == 2 and "cannot be used with" in r.stderr
    r = subprocess.run([BIN, "--quiet", "--verbose", "-qq", "-v", "/nonexistent/x.fa"], capture_output=True, text=True)
    assert r.returncode == 1 and "unexpected argument" not in r.stderr


@pytest.mark.skipif(has_gpu(), reason="needs a box without a GPU")
def test_cli_fails_loudly_without_gpu(tmp_path):
    fa = tmp_path / "r.fa"
    fa.write_text(">a\nACGTACGTAC\n>b\nACGTTTGACA\n>c\nACGTAAAAAA\n")
    r = subprocess.run([BIN, "-n", "3", str(fa)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cli_ava_matches_library(tmp_path):
    g = readsim.genome(120_000)
    R = readsim.reads(g, 150, 3000, readsim.ONT, readsim.SEED_AVA, "r")
    fa = tmp_path / "reads.fa"
    R.write_fasta(str(fa))
    r = subprocess.run([BIN, "-n", "150", "-s", "1", "-f", str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with lrge_b200.Context(0) as ctx:
        want = lrge_b200.AvaStrategy(R, ctx=ctx).estimate(True, lrge_b200.LOWER_QUANTILE, lrge_b200.UPPER_QUANTILE)
    # -f prints the shortest decimal that round-trips the f32 (main.rs:103-107), never an exponent
    assert "e" not in r.stdout and np.float32(r.stdout.strip()) == np.float32(want.estimate)
    assert r.stdout.strip() == np.format_float_positional(np.float32(want.estimate), unique=True, trim="-") or \
        np.float32(r.stdout.strip()) == np.float32(want.estimate)
    assert "Estimated genome size" in r.stderr and "IQR" in r.stderr
    # FASTQ input, two-set, -F, --use-min-ref: runs and lands near the truth
    fq = tmp_path / "reads.fq"
    with open(fq, "wb") as fh:
        for i in range(R.n):
            s = R.seq(i)
            fh.write(b"@" + R.name(i) + b" comment\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    r = subprocess.run([BIN, "-T", "100", "-Q", "50", "-s", "7", "-F", "--use-min-ref", str(fq)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert 40_000 < float(r.stdout.strip()) < 400_000
    # -C -D dir keeps overlaps.paf (twoset.rs:244-250): 17 columns, PafRecord's tags, one row per overlap record
    r = subprocess.run([BIN, "-n", "150", "-s", "1", "-C", "-D", str(tmp_path), str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = open(tmp_path / "overlaps.paf").read().splitlines()
    with lrge_b200.Context(0) as ctx:
        ov, _ = ctx.ava_overlaps(R, "ont")
    assert len(rows) == ov.shape[0] > 100
    assert rows == [lrge_b200.paf_line(o, R, R).decode() for o in ov]
    f = rows[0].split("\t")
    assert len(f) == 17 and f[12] == "tp:A:S" and f[13].startswith("cm:i:") and f[15].startswith("dv:f:") and f[16].startswith("rl:i:")
