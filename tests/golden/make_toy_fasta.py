"""Decode the reference's only real-data fixture, /root/reference/lrge/tests/data/toy.bam (500 unaligned ONT reads,
lrge/tests/alignment.rs:51-68), into tests/golden/toy_reads.fa.gz.

BAM is BGZF (concatenated gzip members, which Python's gzip module reads transparently) around a little-endian
record stream; only the name and the 4-bit packed sequence of each record are needed.  Run in the build container
(the reference tree does not exist on the GPU box); the FASTA is committed."""
import gzip
import os
import struct
import sys

SRC = "/root/reference/lrge/tests/data/toy.bam"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "toy_reads.fa.gz")
CODE = "=ACMGRSVTWYHKDBN"


def main():
    raw = gzip.open(SRC, "rb").read()
    assert raw[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<i", raw, 4)
    p = 8 + l_text
    n_ref, = struct.unpack_from("<i", raw, p)
    p += 4
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", raw, p)
        p += 4 + l_name + 4
    recs = []
    while p < len(raw):
        block_size, = struct.unpack_from("<i", raw, p)
        q = p + 4
        _ref, _pos, l_read_name, _mapq, _bin, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", raw, q)
        q += 32
        name = raw[q:q + l_read_name - 1].decode()
        q += l_read_name + 4 * n_cigar
        packed = raw[q:q + (l_seq + 1) // 2]
        seq = "".join(CODE[b >> 4] + CODE[b & 15] for b in packed)[:l_seq]
        assert flag & 4, "reads are expected to be unmapped"
        recs.append((name, seq))
        p += 4 + block_size
    with gzip.GzipFile(DST, "wb", mtime=0) as f:
        for name, seq in recs:
            f.write(f">{name}\n{seq}\n".encode())
    print(f"{len(recs)} reads, {sum(len(s) for _, s in recs)} bases -> {DST}", file=sys.stderr)


if __name__ == "__main__":
    main()
