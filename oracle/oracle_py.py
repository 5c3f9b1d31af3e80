"""TEST INFRASTRUCTURE — ctypes access to the CPU oracle (oracle/_build/liboracle.so) and, when it
was built, to the reference's own compiled chaining (oracle/_ref/libmm2ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "_build", "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libmm2ref.so")
REF_CLI = os.path.join(HERE, "_ref", "minimap2-sw")
REF_TEST = os.path.join(HERE, "_ref", "test")

ANCHOR = np.dtype([("x", "<u8"), ("y", "<u8")])


class Params(C.Structure):
    """Positional chaining arguments of mm_chain_dp (chain.c:29); defaults = map-ont/asm20 (options.c:24-31)."""
    _fields_ = [(k, C.c_int32) for k in
                ("max_dist_x", "max_dist_y", "bw", "max_skip", "max_iter", "min_cnt", "min_sc", "is_cdna", "n_segs")] + \
               [("gap_scale", C.c_float)]

    def __init__(self, max_dist_x=5000, max_dist_y=5000, bw=500, max_skip=25, max_iter=5000, min_cnt=3, min_sc=40,
                 is_cdna=0, n_segs=1, gap_scale=1.0):
        super().__init__(max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs, gap_scale)

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("cells", "window_cells", "n_anchors", "n_chains", "n_chained")]


REF_FN = C.CFUNCTYPE(C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                     C.c_int64, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_void_p, C.c_int)

_oracle = None
_ref = None


def build(force=False):
    """Compile the oracle (and the in-place reference build when /root/reference exists)."""
    if force or not os.path.exists(LIB_ORACLE) or \
            os.path.getmtime(LIB_ORACLE) < max(os.path.getmtime(os.path.join(HERE, f)) for f in ("chain_oracle.c", "replay.c", "seed_oracle.c", "chain_oracle.h")):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if os.path.exists("/root/reference/chain.c"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "-j8"])


def lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(LIB_ORACLE):
            build()
        L = C.CDLL(LIB_ORACLE)
        L.mm2o_avg_qspan_scaled.restype = C.c_float
        L.mm2o_avg_qspan_scaled.argtypes = [C.c_int64, C.c_void_p]
        L.mm2o_dp_fill.restype = None
        L.mm2o_dp_fill.argtypes = [C.POINTER(Params), C.c_int64] + [C.c_void_p] * 5 + [C.POINTER(Stats)]
        L.mm2o_chain.restype = C.c_int
        L.mm2o_chain.argtypes = [C.POINTER(Params), C.c_int64] + [C.c_void_p] * 4 + \
                                [C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(Stats)]
        L.mm2o_sort_128x.restype = None
        L.mm2o_sort_128x.argtypes = [C.c_void_p, C.c_int64]
        L.mm2o_replay.restype = C.c_double
        L.mm2o_replay.argtypes = [C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(LIB_REF)


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(LIB_REF)
        L.mm_chain_dp_ref.restype = C.c_void_p
        L.mm_chain_dp_ref.argtypes = [C.c_int] * 7 + [C.c_float, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                                     C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
        L.kmalloc.restype = C.c_void_p
        L.kmalloc.argtypes = [C.c_void_p, C.c_size_t]
        L.kfree.restype = None
        L.kfree.argtypes = [C.c_void_p, C.c_void_p]
        L.radix_sort_128x.restype = None
        L.radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
        _ref = L
    return _ref


def _ptr(arr):
    return arr.ctypes.data_as(C.c_void_p) if arr is not None else None


def chain(par, a, want_fpv=False):
    """Oracle for one read. Returns dict(status, u, b, f, p, v, stats)."""
    a = np.ascontiguousarray(a, dtype=ANCHOR)
    n = len(a)
    f = np.empty(n, np.int32) if want_fpv else None
    p = np.empty(n, np.int32) if want_fpv else None
    v = np.empty(n, np.int32) if want_fpv else None
    u = np.empty(max(n, 1), np.uint64)
    b = np.empty(max(n, 1), ANCHOR)
    n_u, n_v, st = C.c_int32(0), C.c_int64(0), Stats()
    status = lib().mm2o_chain(C.byref(par), n, _ptr(a) if n else None, _ptr(f), _ptr(p), _ptr(v),
                              C.byref(n_u), _ptr(u), C.byref(n_v), _ptr(b), C.byref(st))
    return dict(status=status, u=u[:n_u.value].copy(), b=b[:n_v.value].copy(), f=f, p=p, v=v, stats=st)


def ref_chain(par, a):
    """The reference's own compiled mm_chain_dp (software branch) for one read. Returns dict(u, b, b_null, u_null)."""
    L = ref_lib()
    a = np.ascontiguousarray(a, dtype=ANCHOR)
    n = len(a)
    pa = None
    if n:
        pa = L.kmalloc(None, n * 16)      # mm_chain_dp frees it (chain.c:421)
        C.memmove(pa, a.ctypes.data, n * 16)
    n_u, pu = C.c_int(0), C.c_void_p(0)
    pb = L.mm_chain_dp_ref(par.max_dist_x, par.max_dist_y, par.bw, par.max_skip, par.max_iter, par.min_cnt, par.min_sc,
                           par.gap_scale, par.is_cdna, par.n_segs, n, pa, C.byref(n_u), C.byref(pu), None, 0)
    u = np.empty(n_u.value, np.uint64)
    if n_u.value:
        C.memmove(u.ctypes.data, pu.value, n_u.value * 8)
    n_v = int((u & np.uint64(0xffffffff)).sum())
    b = np.empty(n_v, ANCHOR)
    if n_v:
        C.memmove(b.ctypes.data, pb, n_v * 16)
    res = dict(u=u, b=b, b_null=not pb, u_null=not pu.value)
    if pu.value:
        L.kfree(None, pu)
    if pb:
        L.kfree(None, pb)
    return res


def replay(par, off, a, n_threads=1, use_ref=False, want_out=True):
    """Chain a CSR batch on the CPU with n_threads. Returns dict(seconds, n_u, n_v, u, b, stats); u/b at read offsets."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    a = np.ascontiguousarray(a, dtype=ANCHOR)
    n_reads = len(off) - 1
    n_u = np.zeros(n_reads, np.int32)
    n_v = np.zeros(n_reads, np.int64)
    u = np.zeros(max(len(a), 1), np.uint64) if want_out else None
    b = np.zeros(max(len(a), 1), ANCHOR) if want_out else None
    st = Stats()
    fn = None
    if use_ref:
        fn = C.cast(ref_lib().mm_chain_dp_ref, C.c_void_p)
    sec = lib().mm2o_replay(C.byref(par), n_reads, _ptr(off), _ptr(a), int(n_threads), fn, _ptr(n_u), _ptr(n_v), _ptr(u), _ptr(b), C.byref(st))
    return dict(seconds=sec, n_u=n_u, n_v=n_v, u=u, b=b, stats=st)
