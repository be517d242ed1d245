"""Synthetic read sets of the shapes BASELINE.json names. Wraps libreadsim.so."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .readset import ReadSet, csrc_dir

_LIB = None

# error models (per-base independent): (sub, ins, del)
ONT = (0.02, 0.02, 0.02)
HIFI = (0.001, 0.001, 0.001)

SEED_GENOME, SEED_TARGET, SEED_QUERY, SEED_AVA = 42, 43, 44, 45


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(csrc_dir(), "libreadsim.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(path)
        lib.readsim_genome.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
        lib.readsim_genome.restype = None
        lib.readsim_reads.argtypes = [
            ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_double,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint32,
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ]
        lib.readsim_reads.restype = ctypes.c_int
        lib.readsim_free.argtypes = [ctypes.c_void_p]
        lib.readsim_free.restype = None
        _LIB = lib
    return _LIB


def genome(size: int, seed: int = SEED_GENOME) -> np.ndarray:
    g = np.empty(size, dtype=np.uint8)
    _lib().readsim_genome(size, seed, g.ctypes.data)
    return g


def reads(gen: np.ndarray, n: int, mean_len: float, errors=ONT, seed: int = SEED_TARGET,
          prefix: str = "t", n_rate: float = 0.0, first_id: int = 0, truth: bool = False):
    lib = _lib()
    pb, po, pn, pno = (ctypes.c_void_p() for _ in range(4))
    start = np.zeros(n, dtype=np.uint64)
    strand = np.zeros(n, dtype=np.uint8)
    span = np.zeros(n, dtype=np.uint32)
    rc = lib.readsim_reads(gen.ctypes.data, gen.shape[0], n, float(mean_len),
                           errors[0], errors[1], errors[2], n_rate, seed, prefix.encode(), first_id,
                           ctypes.byref(pb), ctypes.byref(po), ctypes.byref(pn), ctypes.byref(pno),
                           start.ctypes.data, strand.ctypes.data, span.ctypes.data)
    if rc != 0:
        raise RuntimeError("readsim_reads failed")
    try:
        off = np.ctypeslib.as_array(ctypes.cast(po, ctypes.POINTER(ctypes.c_uint64)), shape=(n + 1,)).copy()
        noff = np.ctypeslib.as_array(ctypes.cast(pno, ctypes.POINTER(ctypes.c_uint32)), shape=(n + 1,)).copy()
        nb, nn = int(off[-1]), int(noff[-1])
        bases = np.ctypeslib.as_array(ctypes.cast(pb, ctypes.POINTER(ctypes.c_uint8)), shape=(max(nb, 1),))[:nb].copy()
        names = np.ctypeslib.as_array(ctypes.cast(pn, ctypes.POINTER(ctypes.c_uint8)), shape=(max(nn, 1),))[:nn].copy()
    finally:
        for p in (pb, po, pn, pno):
            lib.readsim_free(p)
    rs = ReadSet(bases, off, names, noff)
    if truth:
        return rs, dict(start=start, strand=strand, span=span)
    return rs


# ---- the BASELINE.json configurations -----------------------------------------------------------

CONFIGS = {
    # name: (strategy, n_query, n_target, mean_len, genome_size, platform)
    "C1": ("twoset", 500, 1000, 8000, 5_000_000, "ont"),
    "C2": ("twoset", 5000, 10000, 8000, 4_400_000, "ont"),
    "C3": ("ava", 25000, 0, 15000, 12_000_000, "pb"),
    "C4": ("twoset", 50000, 100000, 10000, 140_000_000, "ont"),
    "C5": ("twoset", 100000, 2000000, 12000, 3_000_000_000, "ont"),
}


def make_config(name: str, scale: float = 1.0, target_mult: int = 1):
    """Returns (strategy, platform, T, Q) for two-set or (strategy, platform, R, None) for all-vs-all.

    scale < 1 shrinks read counts and genome together (coverage preserved) for parity-sized cases;
    target_mult multiplies the target count only (the weak-scaling axis of bench.py --gpus N)."""
    strat, nq, nt, mean_len, gsize, platform = CONFIGS[name]
    nq, nt, gsize = max(1, int(nq * scale)), int(nt * scale) * target_mult, max(20000, int(gsize * scale))
    gen = genome(gsize)
    err = ONT if platform == "ont" else HIFI
    if strat == "twoset":
        T = reads(gen, nt, mean_len, err, SEED_TARGET, "t")
        Q = reads(gen, nq, mean_len, err, SEED_QUERY, "q")
        return strat, platform, T, Q
    R = reads(gen, nq, mean_len, err, SEED_AVA, "r")
    return strat, platform, R, None
