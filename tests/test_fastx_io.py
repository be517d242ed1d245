"""Host-side FASTA / FASTQ ingestion of the library (io.rs:63-205 for the formats in scope: plain and gzip)."""
import gzip

import numpy as np
import pytest

import lrge_b200
from lrge_b200 import readsim
from lrge_b200.readset import ReadSet


@pytest.fixture(scope="module")
def reads():
    g = readsim.genome(30_000)
    return readsim.reads(g, 40, 1500, readsim.ONT, 7, "r")


def same(a: ReadSet, b: ReadSet):
    return (np.array_equal(a.bases, b.bases) and np.array_equal(a.off, b.off) and np.array_equal(a.names, b.names)
            and np.array_equal(a.name_off, b.name_off))


def test_fasta_plain_gz_and_multiline(tmp_path, reads):
    fa = tmp_path / "r.fa"
    reads.write_fasta(str(fa))
    assert same(lrge_b200.read_fastx(str(fa)), reads)
    assert same(ReadSet.from_fasta(str(fa)), reads)
    gz = tmp_path / "r.fa.gz"
    with gzip.open(gz, "wb") as fh:
        fh.write(fa.read_bytes())
    assert same(lrge_b200.read_fastx(str(gz)), reads)
    ml = tmp_path / "multi.fa"                          # wrapped sequence lines, comments after the id, CRLF
    with open(ml, "wb") as fh:
        for i in range(reads.n):
            s = reads.seq(i)
            fh.write(b">" + reads.name(i) + b" some comment\r\n")
            for j in range(0, len(s), 70):
                fh.write(s[j:j + 70] + b"\r\n")
    assert same(lrge_b200.read_fastx(str(ml)), reads)


def test_fastq(tmp_path, reads):
    fq = tmp_path / "r.fq.gz"
    with gzip.open(fq, "wb") as fh:
        for i in range(reads.n):
            s = reads.seq(i)
            fh.write(b"@" + reads.name(i) + b"\tdesc\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    assert same(lrge_b200.read_fastx(str(fq)), reads)


def test_errors(tmp_path):
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b"")
    with pytest.raises(lrge_b200.LrgeError) as e:          # io.rs:140-145 "Is the file empty?"
        lrge_b200.read_fastx(str(empty))
    assert e.value.kind == "IoError"
    with pytest.raises(lrge_b200.LrgeError) as e:
        lrge_b200.read_fastx(str(tmp_path / "missing.fa"))
    assert e.value.kind == "IoError"
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@r1\nACGT\n+\nII\n")                 # quality shorter than sequence
    with pytest.raises(lrge_b200.LrgeError) as e:
        lrge_b200.read_fastx(str(bad))
    assert e.value.kind == "FastqParseError"
    bz = tmp_path / "r.fa.bz2"
    bz.write_bytes(b"BZh91AY&SY")                          # io.rs:47-48 magic bytes
    with pytest.raises(lrge_b200.LrgeError) as e:
        lrge_b200.read_fastx(str(bz))
    assert e.value.kind == "Unsupported"


def test_two_pass_subsampled_ingestion(tmp_path, reads):
    """split_fastq / subsample_reads (twoset.rs:122-201, ava.rs:108-161): pass 1 counts, pass 2 keeps only the drawn
    records, two disjoint sets in one pass, file order inside a set; plain, gzip and FASTQ."""
    fa = tmp_path / "r.fa"
    reads.write_fasta(str(fa))
    gz = tmp_path / "r.fq.gz"
    with gzip.open(gz, "wb") as fh:
        for i in range(reads.n):
            s = reads.seq(i)
            fh.write(b"@" + reads.name(i) + b" x\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    rng = np.random.default_rng(1)
    pick = rng.permutation(reads.n)[:25]
    ia, ib = np.sort(pick[:15]), np.sort(pick[15:])
    for path in (fa, gz):
        A, B = lrge_b200.read_fastx_subsets(str(path), ia, ib)
        assert same(A, reads.subset(ia)) and same(B, reads.subset(ib))
    only = lrge_b200.read_fastx_subsets(str(fa), ia)
    assert same(only, reads.subset(ia))
    for bad in ([3, 3], [5, 2], [reads.n]):             # duplicates, descending, beyond the last record
        with pytest.raises(lrge_b200.LrgeError):
            lrge_b200.read_fastx_subsets(str(fa), bad)


def test_reader_edge_cases(tmp_path):
    """Shapes needletail accepts (io.rs:121-184 reads through it): no trailing newline, blank lines, multi-member gzip,
    lower-case and IUPAC bases kept verbatim, an empty record, '@' as the first quality character."""
    def rd(name, data):
        p = tmp_path / name
        p.write_bytes(data)
        r = lrge_b200.read_fastx(str(p))
        return [(r.name(i), r.seq(i)) for i in range(r.n)]

    assert rd("a.fa", b">r1\nACGT\n>r2 x\nGGTT") == [(b"r1", b"ACGT"), (b"r2", b"GGTT")]
    assert rd("b.fa", b">r1\nACGT\n\n>r2\nGG\n\nTT\n\n") == [(b"r1", b"ACGT"), (b"r2", b"GGTT")]
    assert rd("c.fa.gz", gzip.compress(b">r1\nACGT\n") + gzip.compress(b">r2\nGGTT\n")) == [(b"r1", b"ACGT"), (b"r2", b"GGTT")]
    assert rd("d.fa", b">r1\nacgtNnRy\n") == [(b"r1", b"acgtNnRy")]
    assert rd("e.fa", b">r1\n>r2\nACGT\n") == [(b"r1", b""), (b"r2", b"ACGT")]
    assert rd("f.fq", b"@r1\nACGT\n+r1\nIIII\n@r2\nGG\n+\nII") == [(b"r1", b"ACGT"), (b"r2", b"GG")]
    assert rd("g.fq", b"@r1\nACGT\n+\n@III\n@r2\nGG\n+\n@@\n") == [(b"r1", b"ACGT"), (b"r2", b"GG")]
    with pytest.raises(lrge_b200.LrgeError) as e:
        rd("h.fa", b"hello\nworld\n")
    assert e.value.kind == "FastqParseError"
