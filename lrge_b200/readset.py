"""In-memory read sets in the layout the C-ABI takes (`lrge_readset`, include/lrge_b200.h).

A read set is what liblrge writes to `target.fa` / `query.fa` / `reads.fa` before the hot path
starts (twoset.rs:157-186, ava.rs:139-156): read order in the file is the `rid` order of the index.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np


@dataclass
class ReadSet:
    bases: np.ndarray      # uint8, concatenated ASCII
    off: np.ndarray        # uint64, n+1
    names: np.ndarray      # uint8, concatenated names (no terminators)
    name_off: np.ndarray   # uint32, n+1

    def __post_init__(self):
        self.bases = np.ascontiguousarray(self.bases, dtype=np.uint8)
        self.off = np.ascontiguousarray(self.off, dtype=np.uint64)
        self.names = np.ascontiguousarray(self.names, dtype=np.uint8)
        self.name_off = np.ascontiguousarray(self.name_off, dtype=np.uint32)
        assert self.off.shape[0] == self.name_off.shape[0]

    @property
    def n(self) -> int:
        return int(self.off.shape[0] - 1)

    @property
    def total_bases(self) -> int:
        return int(self.off[-1])

    def lengths(self) -> np.ndarray:
        return np.diff(self.off).astype(np.int64)

    def name(self, i: int) -> bytes:
        return self.names[self.name_off[i]:self.name_off[i + 1]].tobytes()

    def seq(self, i: int) -> bytes:
        return self.bases[int(self.off[i]):int(self.off[i + 1])].tobytes()

    def subset(self, idx) -> "ReadSet":
        idx = np.asarray(idx, dtype=np.int64)
        return ReadSet.from_records([(self.name(int(i)), self.seq(int(i))) for i in idx])

    def concat(self, other: "ReadSet") -> "ReadSet":
        return ReadSet(
            np.concatenate([self.bases, other.bases]),
            np.concatenate([self.off, other.off[1:] + self.off[-1]]),
            np.concatenate([self.names, other.names]),
            np.concatenate([self.name_off, other.name_off[1:] + self.name_off[-1]]),
        )

    @staticmethod
    def from_records(records) -> "ReadSet":
        """records: iterable of (name: bytes|str, seq: bytes|str)."""
        names, seqs = [], []
        for nm, sq in records:
            names.append(nm.encode() if isinstance(nm, str) else bytes(nm))
            seqs.append(sq.encode() if isinstance(sq, str) else bytes(sq))
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        noff = np.zeros(len(seqs) + 1, dtype=np.uint32)
        if seqs:
            off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
            noff[1:] = np.cumsum([len(s) for s in names], dtype=np.uint64)
        return ReadSet(
            np.frombuffer(b"".join(seqs), dtype=np.uint8).copy(),
            off,
            np.frombuffer(b"".join(names), dtype=np.uint8).copy(),
            noff,
        )

    @staticmethod
    def from_fasta(path: str) -> "ReadSet":
        """Plain FASTA/FASTQ; name = header up to the first whitespace (io.rs:192-205, bseq.c)."""
        recs = []
        with open(path, "rb") as fh:
            first = fh.read(1)
            fh.seek(0)
            if first == b"@":
                while True:
                    h = fh.readline()
                    if not h:
                        break
                    s = fh.readline().rstrip(b"\r\n")
                    fh.readline()
                    fh.readline()
                    recs.append((h[1:].split()[0] if h[1:].split() else b"", s))
            else:
                name, chunks = None, []
                for line in fh:
                    if line.startswith(b">"):
                        if name is not None:
                            recs.append((name, b"".join(chunks)))
                        toks = line[1:].split()
                        name, chunks = (toks[0] if toks else b""), []
                    else:
                        chunks.append(line.rstrip(b"\r\n"))
                if name is not None:
                    recs.append((name, b"".join(chunks)))
        return ReadSet.from_records(recs)

    def write_fasta(self, path: str) -> None:
        with open(path, "wb") as fh:
            for i in range(self.n):
                fh.write(b">" + self.name(i) + b"\n" + self.seq(i) + b"\n")


class CReadSet(ctypes.Structure):
    """struct lrge_readset / orc_readset_t (identical layout)."""
    _fields_ = [
        ("bases", ctypes.c_void_p),
        ("off", ctypes.c_void_p),
        ("names", ctypes.c_void_p),
        ("name_off", ctypes.c_void_p),
        ("n", ctypes.c_uint32),
    ]


def as_c(rs: ReadSet) -> CReadSet:
    return CReadSet(rs.bases.ctypes.data, rs.off.ctypes.data, rs.names.ctypes.data, rs.name_off.ctypes.data, rs.n)


def csrc_dir() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
