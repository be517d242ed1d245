"""CPU ORACLE binding (test infrastructure, NOT product code).

ctypes wrapper over oracle/liboracle.so (mm2_oracle.c): a plain-C restatement of the minimap2 v2.30
ava-ont / ava-pb overlap path plus liblrge's counting and estimator logic. PARITY UNPINNED: see the
header of mm2_oracle.h.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; lrge_b200/ never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TIE_KSORT, TIE_STABLE = 0, 1
F_NO_DIAG, F_NO_DUAL = 0x001, 0x002


class Opt(ctypes.Structure):
    _fields_ = [
        ("k", ctypes.c_int32), ("w", ctypes.c_int32), ("is_hpc", ctypes.c_int32), ("bucket_bits", ctypes.c_int32),
        ("flag", ctypes.c_int64),
        ("bw", ctypes.c_int32), ("bw_long", ctypes.c_int32), ("max_gap", ctypes.c_int32), ("max_gap_ref", ctypes.c_int32),
        ("max_chain_skip", ctypes.c_int32), ("max_chain_iter", ctypes.c_int32), ("min_cnt", ctypes.c_int32),
        ("min_chain_score", ctypes.c_int32),
        ("chain_gap_scale", ctypes.c_float), ("chain_skip_scale", ctypes.c_float),
        ("mid_occ_frac", ctypes.c_float), ("q_occ_frac", ctypes.c_float),
        ("min_mid_occ", ctypes.c_int32), ("max_mid_occ", ctypes.c_int32), ("mid_occ", ctypes.c_int32),
        ("max_occ", ctypes.c_int32), ("occ_dist", ctypes.c_int32), ("max_max_occ", ctypes.c_int32),
        ("seed", ctypes.c_int32), ("tie_mode", ctypes.c_int32),
        ("shard_index", ctypes.c_int32), ("shard_count", ctypes.c_int32),
    ]


class Reg(ctypes.Structure):
    _fields_ = [
        ("rid", ctypes.c_int32), ("rev", ctypes.c_int32),
        ("qs", ctypes.c_int32), ("qe", ctypes.c_int32), ("rs", ctypes.c_int32), ("re", ctypes.c_int32),
        ("score", ctypes.c_int32), ("cnt", ctypes.c_int32), ("mlen", ctypes.c_int32), ("blen", ctypes.c_int32),
        ("div", ctypes.c_float), ("hash", ctypes.c_uint32), ("rl", ctypes.c_int32), ("order", ctypes.c_int32),
    ]


REG_DTYPE = np.dtype([("rid", "<i4"), ("rev", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("rs", "<i4"), ("re", "<i4"),
                      ("score", "<i4"), ("cnt", "<i4"), ("mlen", "<i4"), ("blen", "<i4"), ("div", "<f4"), ("hash", "<u4"),
                      ("rl", "<i4"), ("order", "<i4")])
M128_DTYPE = np.dtype([("x", "<u8"), ("y", "<u8")])


class Stats(ctypes.Structure):
    _fields_ = [
        ("n_t_minimizers", ctypes.c_uint64), ("n_q_minimizers", ctypes.c_uint64), ("n_keys", ctypes.c_uint64),
        ("n_seeds", ctypes.c_uint64), ("n_hits", ctypes.c_uint64), ("n_anchors", ctypes.c_uint64),
        ("n_chains", ctypes.c_uint64), ("n_dp_iter", ctypes.c_uint64), ("n_overlaps", ctypes.c_uint64),
        ("no_mapping_count", ctypes.c_uint32), ("mid_occ", ctypes.c_int32),
        ("t_index_s", ctypes.c_double), ("t_map_s", ctypes.c_double),
        ("n_tie_x_queries", ctypes.c_uint64), ("n_tie_x_anchors", ctypes.c_uint64), ("n_mz_flt_deleted", ctypes.c_uint64),
        ("n_max_iter_clamp", ctypes.c_uint64), ("n_skip_break", ctypes.c_uint64), ("n_occ_filtered", ctypes.c_uint64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class CReadSet(ctypes.Structure):
    _fields_ = [("bases", ctypes.c_void_p), ("off", ctypes.c_void_p), ("names", ctypes.c_void_p),
                ("name_off", ctypes.c_void_p), ("n", ctypes.c_uint32)]


def _c_rs(rs) -> CReadSet:
    return CReadSet(rs.bases.ctypes.data, rs.off.ctypes.data, rs.names.ctypes.data, rs.name_off.ctypes.data, rs.n)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "mm2_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, P = ctypes.c_void_p, ctypes.POINTER
        L.orc_opt_init.argtypes = [P(Opt), ctypes.c_char_p, ctypes.c_int]
        L.orc_opt_init.restype = ctypes.c_int
        L.orc_sketch.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32, ctypes.c_int32,
                                 P(vp), P(ctypes.c_int64)]
        L.orc_sketch_readset.argtypes = [P(CReadSet), P(Opt), P(vp), P(ctypes.c_int64), vp]
        L.orc_index_build.argtypes = [P(CReadSet), P(Opt), ctypes.c_int]
        L.orc_index_build.restype = vp
        L.orc_index_destroy.argtypes = [vp]
        L.orc_index_destroy.restype = None
        L.orc_index_n_keys.argtypes = [vp]
        L.orc_index_n_keys.restype = ctypes.c_uint64
        L.orc_index_n_minimizers.argtypes = [vp]
        L.orc_index_n_minimizers.restype = ctypes.c_uint64
        L.orc_index_get.argtypes = [vp, ctypes.c_uint64, P(ctypes.c_int32)]
        L.orc_index_get.restype = P(ctypes.c_uint64)
        L.orc_map.argtypes = [vp, P(Opt), vp, ctypes.c_int32, ctypes.c_char_p, P(vp), P(ctypes.c_int32),
                              P(ctypes.c_int32), P(vp), P(ctypes.c_int64), P(Stats)]
        L.orc_is_internal.argtypes = [ctypes.c_int32] * 7 + [ctypes.c_float]
        L.orc_twoset_counts.argtypes = [P(CReadSet), P(CReadSet), P(Opt), ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                        vp, vp, P(Stats)]
        L.orc_twoset_inverse_counts.argtypes = [P(CReadSet), P(CReadSet), P(Opt), ctypes.c_int, ctypes.c_float,
                                                ctypes.c_int, vp, P(Stats)]
        L.orc_ava_counts.argtypes = [P(CReadSet), P(Opt), ctypes.c_int, ctypes.c_float, ctypes.c_int, vp, P(Stats)]
        L.orc_twoset_regs.argtypes = [P(CReadSet), P(CReadSet), P(Opt), ctypes.c_int, P(vp), P(vp), P(ctypes.c_int64)]
        L.orc_per_read_estimate.argtypes = [ctypes.c_uint64, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
        L.orc_per_read_estimate.restype = ctypes.c_float
        L.orc_median.argtypes = [vp, ctypes.c_int64, ctypes.c_float, ctypes.c_float] + [P(ctypes.c_float)] * 3
        L.orc_estimate.argtypes = [vp, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_float] + [P(ctypes.c_float)] * 3
        L.orc_gap_penalty_table.argtypes = [ctypes.c_int32, ctypes.c_float, ctypes.c_int32, vp]
        L.orc_gap_penalty_table.restype = None
        L.orc_free.argtypes = [vp]
        L.orc_free.restype = None
        _LIB = L
    return _LIB


def _take(ptr, n, dtype):
    """Copy n records out of a malloc'd buffer and free it."""
    n = int(n)
    if not ptr or not ptr.value:
        return np.zeros(0, dtype=dtype)
    buf = (ctypes.c_uint8 * (n * dtype.itemsize)).from_address(ptr.value)
    out = np.frombuffer(buf, dtype=dtype, count=n).copy()
    lib().orc_free(ptr)
    return out


def options(platform: str = "ont", dual: bool = True, tie_mode: int = TIE_KSORT) -> Opt:
    """platform 'ont' -> ava-ont, 'pb' -> ava-pb (twoset.rs:591-594); dual as in aligner.rs:93-103."""
    o = Opt()
    preset = {"ont": b"ava-ont", "nanopore": b"ava-ont", "pb": b"ava-pb", "pacbio": b"ava-pb"}[platform]
    if lib().orc_opt_init(ctypes.byref(o), preset, 1 if dual else 0) != 0:
        raise ValueError(platform)
    o.tie_mode = tie_mode
    return o


def sketch(seq: bytes, w: int, k: int, rid: int = 0, is_hpc: bool = False) -> np.ndarray:
    arr = np.frombuffer(seq, dtype=np.uint8)
    p, n = ctypes.c_void_p(), ctypes.c_int64()
    lib().orc_sketch(arr.ctypes.data, len(seq), w, k, rid, int(is_hpc), ctypes.byref(p), ctypes.byref(n))
    return _take(p, n.value, M128_DTYPE)


def sketch_readset(rs, opt: Opt):
    p, n = ctypes.c_void_p(), ctypes.c_int64()
    cnt = np.zeros(rs.n, dtype=np.uint32)
    c = _c_rs(rs)
    lib().orc_sketch_readset(ctypes.byref(c), ctypes.byref(opt), ctypes.byref(p), ctypes.byref(n), cnt.ctypes.data)
    return _take(p, n.value, M128_DTYPE), cnt


class Index:
    """mm_idx_t stand-in; building it also runs mm_mapopt_update on `opt` (sets opt.mid_occ)."""

    def __init__(self, T, opt: Opt, threads: int = 1):
        self._T = T
        c = _c_rs(T)
        self.opt = opt
        self.h = lib().orc_index_build(ctypes.byref(c), ctypes.byref(opt), threads)

    def close(self):
        if self.h:
            lib().orc_index_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def n_keys(self):
        return lib().orc_index_n_keys(self.h)

    @property
    def n_minimizers(self):
        return lib().orc_index_n_minimizers(self.h)

    def get(self, minier: int) -> np.ndarray:
        n = ctypes.c_int32()
        p = lib().orc_index_get(self.h, minier, ctypes.byref(n))
        return np.array([p[i] for i in range(n.value)], dtype=np.uint64)

    def map(self, seq: bytes, qname: bytes, want_anchors: bool = False):
        arr = np.frombuffer(seq, dtype=np.uint8)
        regs, n_regs, rep = ctypes.c_void_p(), ctypes.c_int32(), ctypes.c_int32()
        anch, n_a = ctypes.c_void_p(), ctypes.c_int64()
        st = Stats()
        rc = lib().orc_map(self.h, ctypes.byref(self.opt), arr.ctypes.data, len(seq), qname,
                           ctypes.byref(regs), ctypes.byref(n_regs), ctypes.byref(rep),
                           ctypes.byref(anch) if want_anchors else None,
                           ctypes.byref(n_a) if want_anchors else None, ctypes.byref(st))
        if rc != 0:
            raise ValueError("Sequence is empty")
        out = dict(regs=_take(regs, n_regs.value, REG_DTYPE), rep_len=rep.value, stats=st.as_dict())
        if want_anchors:
            out["anchors"] = _take(anch, n_a.value, M128_DTYPE)
        return out


def is_internal(qlen, qs, qe, rev, tlen, ts, te, ratio) -> bool:
    return bool(lib().orc_is_internal(qlen, qs, qe, int(rev), tlen, ts, te, ratio))


def twoset_counts(T, Q, platform="ont", remove_internal=False, max_overhang_ratio=0.2, threads=1, tie_mode=TIE_KSORT,
                  shard_index=0, shard_count=1):
    opt = options(platform, dual=True, tie_mode=tie_mode)
    opt.shard_index, opt.shard_count = shard_index, shard_count
    counts = np.zeros(Q.n, dtype=np.uint32)
    had = np.zeros(Q.n, dtype=np.uint8)
    st = Stats()
    ct, cq = _c_rs(T), _c_rs(Q)
    rc = lib().orc_twoset_counts(ctypes.byref(ct), ctypes.byref(cq), ctypes.byref(opt), int(remove_internal),
                                 max_overhang_ratio, threads, counts.ctypes.data, had.ctypes.data, ctypes.byref(st))
    if rc != 0:
        raise RuntimeError(f"oracle twoset failed rc={rc}")
    return counts, had, st.as_dict()


def twoset_inverse_counts(T, Q, platform="ont", remove_internal=False, max_overhang_ratio=0.2, threads=1, tie_mode=TIE_KSORT):
    opt = options(platform, dual=True, tie_mode=tie_mode)
    counts = np.zeros(Q.n, dtype=np.uint32)
    st = Stats()
    ct, cq = _c_rs(T), _c_rs(Q)
    rc = lib().orc_twoset_inverse_counts(ctypes.byref(ct), ctypes.byref(cq), ctypes.byref(opt), int(remove_internal),
                                         max_overhang_ratio, threads, counts.ctypes.data, ctypes.byref(st))
    if rc == -9:
        raise ValueError("DuplicateReadIdentifier")
    if rc != 0:
        raise RuntimeError(f"oracle twoset inverse failed rc={rc}")
    return counts, st.as_dict()


def ava_counts(R, platform="ont", remove_internal=False, max_overhang_ratio=0.2, threads=1, tie_mode=TIE_KSORT):
    opt = options(platform, dual=False, tie_mode=tie_mode)
    counts = np.zeros(R.n, dtype=np.uint32)
    st = Stats()
    cr = _c_rs(R)
    rc = lib().orc_ava_counts(ctypes.byref(cr), ctypes.byref(opt), int(remove_internal), max_overhang_ratio, threads,
                              counts.ctypes.data, ctypes.byref(st))
    if rc == -9:
        raise ValueError("DuplicateReadIdentifier")
    if rc != 0:
        raise RuntimeError(f"oracle ava failed rc={rc}")
    return counts, st.as_dict()


def twoset_regs(T, Q, platform="ont", threads=1, tie_mode=TIE_KSORT, dual=True):
    opt = options(platform, dual=dual, tie_mode=tie_mode)
    regs, qidx, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
    ct, cq = _c_rs(T), _c_rs(Q)
    rc = lib().orc_twoset_regs(ctypes.byref(ct), ctypes.byref(cq), ctypes.byref(opt), threads,
                               ctypes.byref(regs), ctypes.byref(qidx), ctypes.byref(n))
    if rc != 0:
        raise RuntimeError("oracle regs failed")
    return _take(regs, n.value, REG_DTYPE), _take(qidx, n.value, np.dtype("<u4")), opt.mid_occ


def per_read_estimate(read_len, avg_target_len, n_target_reads, n_ovlaps, ovlap_thresh) -> float:
    return float(lib().orc_per_read_estimate(read_len, avg_target_len, n_target_reads, n_ovlaps, ovlap_thresh))


def median(values, lower_q=None, upper_q=None):
    v = np.ascontiguousarray(values, dtype=np.float32)
    lo, md, up = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
    m = lib().orc_median(v.ctypes.data, v.shape[0], -1.0 if lower_q is None else lower_q,
                         -1.0 if upper_q is None else upper_q, ctypes.byref(lo), ctypes.byref(md), ctypes.byref(up))
    return (lo.value if m & 1 else None, md.value if m & 2 else None, up.value if m & 4 else None)


def estimate(estimates, finite=True, lower_q=None, upper_q=None):
    v = np.ascontiguousarray(estimates, dtype=np.float32)
    lo, md, up = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
    m = lib().orc_estimate(v.ctypes.data, v.shape[0], int(finite), -1.0 if lower_q is None else lower_q,
                           -1.0 if upper_q is None else upper_q, ctypes.byref(lo), ctypes.byref(md), ctypes.byref(up))
    return (lo.value if m & 1 else None, md.value if m & 2 else None, up.value if m & 4 else None)


def gap_penalty_table(k: int, n: int, chain_gap_scale: float = 0.8) -> np.ndarray:
    out = np.zeros(n, dtype=np.int32)
    lib().orc_gap_penalty_table(k, chain_gap_scale, n, out.ctypes.data)
    return out
