"""lrge_b200 -- ctypes glue over liblrge_b200.so (the sm_100a CUDA implementation of liblrge's overlap hot path).

This package is test / bench plumbing around the C-ABI in include/lrge_b200.h; the product is the shared
library.  The class names mirror liblrge's surface (`TwoSetStrategy`, `AvaStrategy`, `Estimate.estimate`,
`EstimateResult`; liblrge/src/{twoset,ava,estimate}.rs) so parity tests read like the reference's own.
There is no CPU fallback: creating a `Context` without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .readset import CReadSet, ReadSet, as_c, csrc_dir

LOWER_QUANTILE = 0.15   # estimate.rs:4
UPPER_QUANTILE = 0.65   # estimate.rs:6
DEFAULT_TARGET_NUM_READS = 10_000   # twoset.rs:65
DEFAULT_QUERY_NUM_READS = 5_000     # twoset.rs:66
DEFAULT_AVA_NUM_READS = 25_000      # ava.rs:62

LRGE_OK = 0
STATUS = {
    0: "ok", 1: "IoError", 2: "FastqParseError", 3: "TooManyReadsError", 4: "TooFewReadsError", 5: "InvalidPlatform",
    6: "ThreadError", 7: "PafWriteError", 8: "MapError", 9: "DuplicateReadIdentifier",
    100: "CudaError", 101: "NcclError", 102: "InvalidArgument", 103: "OutOfMemory", 104: "Unsupported",
}

C_API = [
    "lrge_opts_default", "lrge_gpu_create", "lrge_gpu_destroy", "lrge_gpu_last_error",
    "lrge_gpu_twoset_counts", "lrge_gpu_ava_counts", "lrge_gpu_stage", "lrge_gpu_unstage",
    "lrge_gpu_twoset_counts_staged", "lrge_gpu_ava_counts_staged", "lrge_gpu_host_alloc", "lrge_gpu_host_free",
    "lrge_gpu_dump_minimizers", "lrge_gpu_dump_mid_occ", "lrge_per_read_estimate", "lrge_estimate", "lrge_is_internal",
    "lrge_twoset_generate_estimates", "lrge_ava_generate_estimates", "lrge_gpu_comm_unique_id", "lrge_gpu_comm_init", "lrge_gpu_twoset_counts_shard",
    "lrge_gpu_set_profiling", "lrge_gpu_kernel_times", "lrge_gpu_twoset_overlaps", "lrge_gpu_ava_overlaps", "lrge_paf_format",
    "lrge_gpu_dump_minimizers_block", "lrge_fastx_count", "lrge_fastx_load",
    "lrge_fastx_load_subsets", "lrge_hostset_view", "lrge_hostset_free",
]


class LrgeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code
        self.kind = STATUS.get(code, str(code))


class Opts(ctypes.Structure):
    _fields_ = [
        ("platform", ctypes.c_int32), ("remove_internal", ctypes.c_int32), ("max_overhang_ratio", ctypes.c_float),
        ("use_min_ref", ctypes.c_int32), ("shard_index", ctypes.c_int32), ("shard_count", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 6),
    ]


OVERLAP_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("rev", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("ts", "<i4"),
                          ("te", "<i4"), ("score", "<i4"), ("cnt", "<i4"), ("mlen", "<i4"), ("blen", "<i4"), ("dv", "<f4"), ("rl", "<i4")])


class KernelTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 64), ("ms", ctypes.c_double), ("algorithmic_bytes", ctypes.c_double),
                ("launches", ctypes.c_uint32), ("passes", ctypes.c_uint32)]


class Stats(ctypes.Structure):
    _fields_ = [
        ("n_target_bases", ctypes.c_uint64), ("n_query_bases", ctypes.c_uint64),
        ("n_target_minimizers", ctypes.c_uint64), ("n_query_minimizers", ctypes.c_uint64), ("n_keys", ctypes.c_uint64),
        ("n_seeds", ctypes.c_uint64), ("n_hits", ctypes.c_uint64), ("n_anchors", ctypes.c_uint64),
        ("n_pairs", ctypes.c_uint64), ("n_overlaps", ctypes.c_uint64),
        ("no_mapping_count", ctypes.c_uint32), ("mid_occ", ctypes.c_int32),
        ("n_batches", ctypes.c_uint32), ("n_kernel_launches", ctypes.c_uint32),
        ("ms_h2d", ctypes.c_float), ("ms_pack", ctypes.c_float), ("ms_sketch", ctypes.c_float),
        ("ms_index", ctypes.c_float), ("ms_seed", ctypes.c_float), ("ms_sort", ctypes.c_float),
        ("ms_chain", ctypes.c_float), ("ms_d2h", ctypes.c_float), ("ms_total", ctypes.c_float),
        ("ms_chain_kernel", ctypes.c_float), ("ms_comm", ctypes.c_float),
        ("n_deferred", ctypes.c_uint32), ("n_rounds", ctypes.c_uint32), ("n_index_runs", ctypes.c_uint32),
        ("reserved", ctypes.c_float * 2),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if f != "reserved"}


_LIB = None


def library_path() -> str:
    # LRGE_B200_LIB: tests load the test-only build (liblrge_b200_v1check.so) in a subprocess
    return os.environ.get("LRGE_B200_LIB") or os.path.join(csrc_dir(), "liblrge_b200.so")


def lib():
    """Load liblrge_b200.so; fails loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: the CUDA extension is not built (run __graft_entry__.build()); "
                               "lrge_b200 has no CPU fallback")
        L = ctypes.CDLL(path)
        vp, P = ctypes.c_void_p, ctypes.POINTER
        L.lrge_opts_default.argtypes = [P(Opts)]
        L.lrge_opts_default.restype = None
        L.lrge_gpu_create.argtypes = [ctypes.c_int, P(vp)]
        L.lrge_gpu_destroy.argtypes = [vp]
        L.lrge_gpu_destroy.restype = None
        L.lrge_gpu_last_error.argtypes = [vp]
        L.lrge_gpu_last_error.restype = ctypes.c_char_p
        L.lrge_gpu_twoset_counts.argtypes = [vp, P(CReadSet), P(CReadSet), P(Opts), vp, vp, P(Stats)]
        L.lrge_gpu_twoset_counts_shard.argtypes = [vp, P(CReadSet), P(CReadSet), P(Opts), vp, vp, P(Stats)]
        L.lrge_gpu_ava_counts.argtypes = [vp, P(CReadSet), P(Opts), vp, P(Stats)]
        L.lrge_gpu_stage.argtypes = [vp, P(CReadSet), P(vp)]
        L.lrge_gpu_unstage.argtypes = [vp, vp]
        L.lrge_gpu_unstage.restype = None
        L.lrge_gpu_twoset_counts_staged.argtypes = [vp, vp, vp, P(Opts), vp, vp, P(Stats)]
        L.lrge_gpu_ava_counts_staged.argtypes = [vp, vp, P(Opts), vp, P(Stats)]
        L.lrge_gpu_comm_unique_id.argtypes = [vp, ctypes.c_size_t]
        L.lrge_gpu_comm_init.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_size_t]
        L.lrge_gpu_set_profiling.argtypes = [vp, ctypes.c_int]
        L.lrge_gpu_kernel_times.argtypes = [vp, P(KernelTime), ctypes.c_int]
        L.lrge_gpu_twoset_overlaps.argtypes = [vp, P(CReadSet), P(CReadSet), P(Opts), vp, ctypes.c_uint64, P(ctypes.c_uint64), vp, vp, P(Stats)]
        L.lrge_gpu_ava_overlaps.argtypes = [vp, P(CReadSet), P(Opts), vp, ctypes.c_uint64, P(ctypes.c_uint64), vp, P(Stats)]
        L.lrge_paf_format.argtypes = [vp, ctypes.c_char_p, ctypes.c_int, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int, ctypes.c_int32,
                                      ctypes.c_float, ctypes.c_int32, ctypes.c_char_p, ctypes.c_size_t]
        L.lrge_fastx_count.argtypes = [ctypes.c_char_p, P(ctypes.c_uint64), P(ctypes.c_uint64), P(ctypes.c_uint64)]
        L.lrge_fastx_load.argtypes = [ctypes.c_char_p, vp, ctypes.c_uint64, vp, ctypes.c_uint64, vp, ctypes.c_uint64, vp]
        L.lrge_fastx_load_subsets.argtypes = [ctypes.c_char_p, vp, ctypes.c_uint64, vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, P(vp), P(vp)]
        L.lrge_hostset_view.argtypes = [vp]
        L.lrge_hostset_view.restype = P(CReadSet)
        L.lrge_hostset_free.argtypes = [vp]
        L.lrge_hostset_free.restype = None
        L.lrge_gpu_host_alloc.argtypes = [ctypes.c_size_t, P(vp)]
        L.lrge_gpu_host_free.argtypes = [vp]
        L.lrge_gpu_host_free.restype = None
        L.lrge_gpu_dump_minimizers.argtypes = [vp, P(CReadSet), ctypes.c_int32, vp, vp, ctypes.c_uint64,
                                               P(ctypes.c_uint64), vp]
        L.lrge_gpu_dump_minimizers_block.argtypes = [vp, P(CReadSet), ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32, vp, vp,
                                                     ctypes.c_uint64, P(ctypes.c_uint64), vp]
        L.lrge_gpu_dump_mid_occ.argtypes = [vp, P(CReadSet), ctypes.c_int32, P(ctypes.c_int32), P(ctypes.c_uint64)]
        L.lrge_per_read_estimate.argtypes = [ctypes.c_uint64, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
        L.lrge_per_read_estimate.restype = ctypes.c_float
        L.lrge_estimate.argtypes = [vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_float, ctypes.c_float] + [P(ctypes.c_float)] * 3
        L.lrge_is_internal.argtypes = [ctypes.c_int32] * 7 + [ctypes.c_float]
        L.lrge_twoset_generate_estimates.argtypes = [vp, P(CReadSet), P(CReadSet), P(Opts), vp, P(ctypes.c_uint32), P(Stats)]
        L.lrge_ava_generate_estimates.argtypes = [vp, P(CReadSet), P(Opts), vp, P(ctypes.c_uint32), P(Stats)]
        _LIB = L
    return _LIB


def platform_code(platform) -> int:
    """Platform::from_str (lib.rs:170-180)."""
    if isinstance(platform, int):
        return platform
    s = str(platform).lower()
    if s in ("nanopore", "ont"):
        return 0
    if s in ("pacbio", "pb"):
        return 1
    raise LrgeError(5, str(platform))


def make_opts(platform="ont", remove_internal=False, max_overhang_ratio=0.2, use_min_ref=False,
              shard_index=0, shard_count=1) -> Opts:
    o = Opts()
    lib().lrge_opts_default(ctypes.byref(o))
    o.platform = platform_code(platform)
    o.remove_internal = int(remove_internal)
    o.max_overhang_ratio = max_overhang_ratio
    o.use_min_ref = int(use_min_ref)
    o.shard_index, o.shard_count = shard_index, shard_count
    return o


@dataclass
class Counts:
    counts: np.ndarray
    had_mapping: Optional[np.ndarray]
    stats: dict


class PinnedBuffer:
    """Pinned host memory from the library (for the e2e leg of bench.py)."""

    def __init__(self, nbytes: int):
        self.ptr = ctypes.c_void_p()
        rc = lib().lrge_gpu_host_alloc(max(nbytes, 1), ctypes.byref(self.ptr))
        if rc != 0:
            raise LrgeError(rc, "cudaHostAlloc failed")
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(ctypes.cast(self.ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(max(nbytes, 1),))[:nbytes]

    def free(self):
        if self.ptr:
            lib().lrge_gpu_host_free(self.ptr)
            self.ptr = None


def pinned_copy(rs: ReadSet):
    """A ReadSet whose base array lives in pinned host memory. Returns (readset, buffer-to-free)."""
    buf = PinnedBuffer(rs.bases.shape[0])
    buf.array[:] = rs.bases
    out = ReadSet.__new__(ReadSet)
    out.bases, out.off, out.names, out.name_off = buf.array, rs.off, rs.names, rs.name_off
    return out, buf


class Staged:
    def __init__(self, ctx: "Context", rs: ReadSet):
        self.ctx, self.rs = ctx, rs
        self.h = ctypes.c_void_p()
        c = as_c(rs)
        ctx._check(lib().lrge_gpu_stage(ctx.h, ctypes.byref(c), ctypes.byref(self.h)))

    def free(self):
        if self.h:
            lib().lrge_gpu_unstage(self.ctx.h, self.h)
            self.h = None


class Context:
    """lrge_ctx: one per process per GPU."""

    def __init__(self, device: int = 0):
        self.h = ctypes.c_void_p()
        rc = lib().lrge_gpu_create(device, ctypes.byref(self.h))
        if rc != 0:
            raise LrgeError(rc, lib().lrge_gpu_last_error(None).decode())

    def close(self):
        if self.h:
            lib().lrge_gpu_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise LrgeError(rc, lib().lrge_gpu_last_error(self.h).decode())

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (ctypes.c_uint8 * 128)()
        rc = lib().lrge_gpu_comm_unique_id(buf, 128)
        if rc != 0:
            raise LrgeError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(lib().lrge_gpu_comm_init(self.h, rank, world, buf, 128))
        self.rank, self.world = rank, world

    def set_profiling(self, on: bool):
        self._check(lib().lrge_gpu_set_profiling(self.h, int(on)))

    def kernel_times(self):
        """Per-kernel device time of the last *_counts call (needs set_profiling(True))."""
        buf = (KernelTime * 64)()
        n = lib().lrge_gpu_kernel_times(self.h, buf, 64)
        return [dict(name=buf[i].name.decode(), ms=buf[i].ms, algorithmic_bytes=buf[i].algorithmic_bytes, launches=buf[i].launches, passes=max(1, buf[i].passes))
                for i in range(max(0, min(n, 64)))]

    def stage(self, rs: ReadSet) -> Staged:
        return Staged(self, rs)

    def twoset_counts(self, T, Q, platform="ont", **kw) -> Counts:
        o = make_opts(platform, **kw)
        st = Stats()
        if isinstance(T, Staged):
            nq = Q.rs.n
            counts, had = np.zeros(nq, np.uint32), np.zeros(nq, np.uint8)
            self._check(lib().lrge_gpu_twoset_counts_staged(self.h, T.h, Q.h, ctypes.byref(o), counts.ctypes.data,
                                                            had.ctypes.data, ctypes.byref(st)))
        else:
            counts, had = np.zeros(Q.n, np.uint32), np.zeros(Q.n, np.uint8)
            ct, cq = as_c(T), as_c(Q)
            self._check(lib().lrge_gpu_twoset_counts(self.h, ctypes.byref(ct), ctypes.byref(cq), ctypes.byref(o),
                                                     counts.ctypes.data, had.ctypes.data, ctypes.byref(st)))
        return Counts(counts, had, st.as_dict())

    def twoset_counts_shard(self, T_block, Q, platform="ont", **kw) -> Counts:
        """lrge_gpu_twoset_counts_shard: this rank's block of the targets + all queries (needs comm_init)."""
        o = make_opts(platform, **kw)
        st = Stats()
        counts, had = np.zeros(Q.n, np.uint32), np.zeros(Q.n, np.uint8)
        ct, cq = as_c(T_block), as_c(Q)
        self._check(lib().lrge_gpu_twoset_counts_shard(self.h, ctypes.byref(ct), ctypes.byref(cq), ctypes.byref(o),
                                                       counts.ctypes.data, had.ctypes.data, ctypes.byref(st)))
        return Counts(counts, had, st.as_dict())

    def ava_counts(self, R, platform="ont", **kw) -> Counts:
        o = make_opts(platform, **kw)
        st = Stats()
        if isinstance(R, Staged):
            counts = np.zeros(R.rs.n, np.uint32)
            self._check(lib().lrge_gpu_ava_counts_staged(self.h, R.h, ctypes.byref(o), counts.ctypes.data, ctypes.byref(st)))
        else:
            counts = np.zeros(R.n, np.uint32)
            cr = as_c(R)
            self._check(lib().lrge_gpu_ava_counts(self.h, ctypes.byref(cr), ctypes.byref(o), counts.ctypes.data, ctypes.byref(st)))
        return Counts(counts, None, st.as_dict())

    def twoset_overlaps(self, T: ReadSet, Q: ReadSet, platform="ont", cap: int = 0, **kw):
        """Every overlap (kept chain) as a structured array (OVERLAP_DTYPE) plus the Counts of the same run."""
        o = make_opts(platform, **kw)
        cap = cap or max(1024, 64 * Q.n)
        while True:
            out = np.zeros(cap, OVERLAP_DTYPE)
            n = ctypes.c_uint64()
            counts, had = np.zeros(Q.n, np.uint32), np.zeros(Q.n, np.uint8)
            st = Stats()
            ct, cq = as_c(T), as_c(Q)
            self._check(lib().lrge_gpu_twoset_overlaps(self.h, ctypes.byref(ct), ctypes.byref(cq), ctypes.byref(o), out.ctypes.data, cap,
                                                       ctypes.byref(n), counts.ctypes.data, had.ctypes.data, ctypes.byref(st)))
            if n.value <= cap:
                return out[:n.value], Counts(counts, had, st.as_dict())
            cap = int(n.value)

    def ava_overlaps(self, R: ReadSet, platform="ont", cap: int = 0, **kw):
        o = make_opts(platform, **kw)
        cap = cap or max(1024, 64 * R.n)
        while True:
            out = np.zeros(cap, OVERLAP_DTYPE)
            n = ctypes.c_uint64()
            counts = np.zeros(R.n, np.uint32)
            st = Stats()
            cr = as_c(R)
            self._check(lib().lrge_gpu_ava_overlaps(self.h, ctypes.byref(cr), ctypes.byref(o), out.ctypes.data, cap, ctypes.byref(n),
                                                    counts.ctypes.data, ctypes.byref(st)))
            if n.value <= cap:
                return out[:n.value], Counts(counts, None, st.as_dict())
            cap = int(n.value)

    def dump_minimizers(self, rs: ReadSet, platform="ont"):
        cap = rs.total_bases + 16
        x, y = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        cnt = np.zeros(rs.n, np.uint32)
        n = ctypes.c_uint64()
        c = as_c(rs)
        self._check(lib().lrge_gpu_dump_minimizers(self.h, ctypes.byref(c), platform_code(platform), x.ctypes.data,
                                                   y.ctypes.data, cap, ctypes.byref(n), cnt.ctypes.data))
        return x[:n.value].copy(), y[:n.value].copy(), cnt

    def dump_minimizers_block(self, rs: ReadSet, first: int, count: int, platform="ont"):
        cap = rs.total_bases + 16
        x, y = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        cnt = np.zeros(count, np.uint32)
        n = ctypes.c_uint64()
        c = as_c(rs)
        self._check(lib().lrge_gpu_dump_minimizers_block(self.h, ctypes.byref(c), platform_code(platform), first, count,
                                                         x.ctypes.data, y.ctypes.data, cap, ctypes.byref(n), cnt.ctypes.data))
        return x[:n.value].copy(), y[:n.value].copy(), cnt

    def dump_mid_occ(self, T: ReadSet, platform="ont"):
        m, nk = ctypes.c_int32(), ctypes.c_uint64()
        c = as_c(T)
        self._check(lib().lrge_gpu_dump_mid_occ(self.h, ctypes.byref(c), platform_code(platform), ctypes.byref(m), ctypes.byref(nk)))
        return m.value, nk.value


def read_fastx(path: str) -> ReadSet:
    """Plain or gzip FASTA / FASTQ through the library's host reader (lrge_fastx_count / lrge_fastx_load)."""
    n, nb, nn = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    rc = lib().lrge_fastx_count(path.encode(), ctypes.byref(n), ctypes.byref(nb), ctypes.byref(nn))
    if rc != 0:
        raise LrgeError(rc, path)
    bases, off = np.zeros(max(nb.value, 1), np.uint8), np.zeros(n.value + 1, np.uint64)
    names, noff = np.zeros(max(nn.value, 1), np.uint8), np.zeros(n.value + 1, np.uint32)
    rc = lib().lrge_fastx_load(path.encode(), bases.ctypes.data, nb.value, off.ctypes.data, n.value, names.ctypes.data, nn.value, noff.ctypes.data)
    if rc != 0:
        raise LrgeError(rc, path)
    return ReadSet(bases[:nb.value], off, names[:nn.value], noff)


def read_fastx_subsets(path: str, idx_a, idx_b=None, pinned: bool = False):
    """Pass 2 of the two-pass subsampled ingestion (lrge_fastx_load_subsets): the records idx_a (and idx_b) of the file, in
    file order, as ReadSets (copied out of the library's buffers)."""
    a = np.ascontiguousarray(idx_a, dtype=np.uint64)
    b = np.ascontiguousarray(idx_b if idx_b is not None else [], dtype=np.uint64)
    ha, hb = ctypes.c_void_p(), ctypes.c_void_p()
    rc = lib().lrge_fastx_load_subsets(path.encode(), a.ctypes.data, a.shape[0], b.ctypes.data if b.shape[0] else None, b.shape[0], 0,
                                       int(pinned), ctypes.byref(ha), ctypes.byref(hb) if idx_b is not None else None)
    if rc != 0:
        raise LrgeError(rc, path)

    def take(h):
        v = lib().lrge_hostset_view(h).contents
        n = v.n
        off = np.ctypeslib.as_array(ctypes.cast(v.off, ctypes.POINTER(ctypes.c_uint64)), shape=(n + 1,)).copy()
        noff = np.ctypeslib.as_array(ctypes.cast(v.name_off, ctypes.POINTER(ctypes.c_uint32)), shape=(n + 1,)).copy()
        bases = np.ctypeslib.as_array(ctypes.cast(v.bases, ctypes.POINTER(ctypes.c_uint8)), shape=(max(int(off[-1]), 1),))[:int(off[-1])].copy()
        names = np.ctypeslib.as_array(ctypes.cast(v.names, ctypes.POINTER(ctypes.c_uint8)), shape=(max(int(noff[-1]), 1),))[:int(noff[-1])].copy()
        lib().lrge_hostset_free(h)
        return ReadSet(bases, off, names, noff)

    A = take(ha)
    return (A, take(hb)) if idx_b is not None else A


def paf_line(ov, Q: ReadSet, T: ReadSet, dv=None, rl=None) -> bytes:
    """One PAF row as PafRecord serialises it (mapping.rs:81-177); dv / rl default to the record's own."""
    rec = np.zeros(1, OVERLAP_DTYPE)
    rec[0] = ov
    dv = float(ov["dv"]) if dv is None else dv
    rl = int(ov["rl"]) if rl is None else rl
    qn, tn = Q.name(int(ov["query"])), T.name(int(ov["target"]))
    buf = ctypes.create_string_buffer(1024)
    n = lib().lrge_paf_format(rec.ctypes.data, qn, len(qn), int(Q.off[int(ov["query"]) + 1] - Q.off[int(ov["query"])]), tn, len(tn),
                              int(T.off[int(ov["target"]) + 1] - T.off[int(ov["target"])]), dv, rl, buf, 1024)
    return buf.raw[:n]


# ---- host-side mirror of liblrge's estimator surface -----------------------------------------------

def per_read_estimate(read_len, avg_target_len, n_target_reads, n_ovlaps, ovlap_thresh) -> float:
    return float(lib().lrge_per_read_estimate(read_len, avg_target_len, n_target_reads, n_ovlaps, ovlap_thresh))


def is_internal(qlen, qs, qe, rev, tlen, ts, te, ratio) -> bool:
    return bool(lib().lrge_is_internal(qlen, qs, qe, int(rev), tlen, ts, te, ratio))


@dataclass
class EstimateResult:   # estimate.rs:8-17
    lower: Optional[float]
    estimate: Optional[float]
    upper: Optional[float]
    no_mapping_count: int


def estimate_from(estimates, finite=True, lower_quant=None, upper_quant=None):
    v = np.ascontiguousarray(estimates, dtype=np.float32)
    lo, md, up = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
    m = lib().lrge_estimate(v.ctypes.data, v.shape[0], int(finite), -1.0 if lower_quant is None else lower_quant,
                            -1.0 if upper_quant is None else upper_quant, ctypes.byref(lo), ctypes.byref(md), ctypes.byref(up))
    if m < 0:
        raise ValueError("Quantile must be between 0.0 and 1.0")
    return (lo.value if m & 1 else None, md.value if m & 2 else None, up.value if m & 4 else None)


class _Estimate:
    """trait Estimate (estimate.rs:21-78)."""

    def generate_estimates(self):
        raise NotImplementedError

    def estimate(self, finite: bool = True, lower_quant: Optional[float] = None, upper_quant: Optional[float] = None) -> EstimateResult:
        estimates, no_mapping_count = self.generate_estimates()
        lo, md, up = estimate_from(estimates, finite, lower_quant, upper_quant)
        return EstimateResult(lo, md, up, no_mapping_count)


class TwoSetStrategy(_Estimate):
    """TwoSetStrategy over explicit target / query read sets (twoset.rs:73-99, 587-606; the random
    split of twoset.rs:122-201 happens before the hot path and is the caller's)."""

    def __init__(self, target: ReadSet, query: ReadSet, platform="ont", remove_internal=False, max_overhang_ratio=0.2,
                 use_min_ref=False, device: int = 0, ctx: Optional[Context] = None):
        self.target, self.query = target, query
        self.opts = dict(platform=platform, remove_internal=remove_internal, max_overhang_ratio=max_overhang_ratio,
                         use_min_ref=use_min_ref)
        self.device, self.ctx = device, ctx
        self.stats = None

    def generate_estimates(self):
        ctx = self.ctx or Context(self.device)
        try:
            o = make_opts(**self.opts)
            est = np.zeros(self.query.n, np.float32)
            nomap = ctypes.c_uint32()
            st = Stats()
            ct, cq = as_c(self.target), as_c(self.query)
            ctx._check(lib().lrge_twoset_generate_estimates(ctx.h, ctypes.byref(ct), ctypes.byref(cq), ctypes.byref(o),
                                                            est.ctypes.data, ctypes.byref(nomap), ctypes.byref(st)))
            self.stats = st.as_dict()
            return est, nomap.value
        finally:
            if self.ctx is None:
                ctx.close()


class AvaStrategy(_Estimate):
    """AvaStrategy over an explicit read set (ava.rs:66-84, 369-381)."""

    def __init__(self, reads: ReadSet, platform="ont", remove_internal=False, max_overhang_ratio=0.2, device: int = 0,
                 ctx: Optional[Context] = None):
        self.reads = reads
        self.opts = dict(platform=platform, remove_internal=remove_internal, max_overhang_ratio=max_overhang_ratio)
        self.device, self.ctx = device, ctx
        self.stats = None

    def generate_estimates(self):
        ctx = self.ctx or Context(self.device)
        try:
            o = make_opts(**self.opts)
            est = np.zeros(self.reads.n, np.float32)
            nomap = ctypes.c_uint32()
            st = Stats()
            cr = as_c(self.reads)
            ctx._check(lib().lrge_ava_generate_estimates(ctx.h, ctypes.byref(cr), ctypes.byref(o), est.ctypes.data,
                                                         ctypes.byref(nomap), ctypes.byref(st)))
            self.stats = st.as_dict()
            return est, nomap.value
        finally:
            if self.ctx is None:
                ctx.close()
