"""ctypes view of the C ABI in include/mm2chain_b200.h (libmm2chain_b200.so, built in-tree by build.py).

This is the Python mirror of the reference's accelerator boundary (chain_hardware.h:68-72 / chain.c:29):
  init()/shutdown()      <- hardware_init()/cleanup()
  chain_batch()          <- run_chaining_on_hw(), but many reads per call and final chains out
  chain_read()           <- mm_chain_dp() itself (same positional arguments)
  DeviceBatch            <- inputs/outputs resident in HBM (torch tensors only as device memory)
There is NO CPU fallback: if the library is missing, or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MM2B_LIB") or os.path.join(HERE, "libmm2chain_b200.so")      # MM2B_LIB: tuning experiments only
ANCHOR = np.dtype([("x", "<u8"), ("y", "<u8")])
READ_EMPTY, READ_NO_CHAIN, READ_OK = 0, 1, 2

EXPORTS = ["mm2b_init", "mm2b_init_async", "mm2b_shutdown", "mm2b_shutdown_at_exit", "mm2b_num_devices", "mm2b_cuda_device_count", "mm2b_last_error", "mm2b_abi_version", "mm2b_ws_set_longest_read",
           "mm2b_host_alloc", "mm2b_host_free", "mm2b_host_reserve", "mm2b_host_pool_trim", "mm2b_reserve_for_mapping", "mm2b_chain_batch", "mm2b_ws_create", "mm2b_ws_destroy", "mm2b_ws_bytes",
           "mm2b_ws_set_counting", "mm2b_set_counting",
           "mm2b_chain_batch_device", "mm2b_chain_batch_device_idx", "mm2b_chain_batch_ex", "mm2b_unpack_anchors_device", "mm2b_pack_anchors", "mm2b_measure_host_copy", "mm2b_ws_stats", "mm2b_ws_chain_kernel_ms", "mm2b_launch_count", "mm2b_ws_copy_fpv", "mm2b_debug_flags", "mm2b_measure_int32_peak",
           "mm_chain_dp",
           # include/mm2seed_b200.h: the seeding front end
           "mm2b_index_create", "mm2b_index_destroy", "mm2b_index_lookup", "mm2b_map_supported", "mm2b_map_batch", "mm2b_map_result_release",
           "mm2b_seed_debug", "mm2b_free"]


class Params(C.Structure):
    """mm2b_params_t — the chaining arguments of mm_chain_dp (chain.c:29); defaults are map-ont / asm20 (options.c:24-31)."""
    _fields_ = [(k, C.c_int32) for k in
                ("max_dist_x", "max_dist_y", "bw", "max_skip", "max_iter", "min_cnt", "min_sc", "is_cdna", "n_segs")] + \
               [("gap_scale", C.c_float)]

    def __init__(self, max_dist_x=5000, max_dist_y=5000, bw=500, max_skip=25, max_iter=5000, min_cnt=3, min_sc=40,
                 is_cdna=0, n_segs=1, gap_scale=1.0):
        super().__init__(max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs, gap_scale)

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_reads", "n_anchors", "n_chains", "n_chained", "cells_issued", "cells_ref", "window_cells", "n_general_reads")] + \
               [(k, C.c_double) for k in ("h2d_ms", "kernel_ms", "d2h_ms")] + [("n_heavy_reads", C.c_int64)] + \
               [(k, C.c_int64) for k in ("h2d_bytes", "d2h_bytes", "n_packed_subs", "n_raw_subs")] + [(k, C.c_double) for k in ("pack_ms", "gather_ms")] + [("n_cut_reads", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class IndexDesc(C.Structure):
    """mm2b_index_desc_t (include/mm2seed_b200.h): the reference's minimizer index as flat arrays."""
    _fields_ = [(k, C.c_int32) for k in ("k", "w", "is_hpc", "n_seq")] + [("n_keys", C.c_int64), ("n_pos", C.c_int64)] + [(k, C.c_void_p) for k in ("keys", "vals", "pos")]


class SeedParams(C.Structure):
    _fields_ = [("max_occ", C.c_int32), ("reserved", C.c_int32)]


class MapResult(C.Structure):
    """mm2b_map_result_t"""
    _fields_ = [("n_reads", C.c_int64)] + [(k, C.POINTER(C.c_int32)) for k in ("status", "n_u", "n_v", "rep_len", "n_mini_pos", "n_mini", "seg")] + \
               [(k, C.POINTER(C.c_int64)) for k in ("n_a", "u_off", "b_off", "mp_off")] + [("n_segs", C.c_int32)] + \
               [("seg_u", C.POINTER(C.c_void_p)), ("seg_b", C.POINTER(C.c_void_p)), ("seg_mini_pos", C.POINTER(C.c_void_p))] + \
               [(k, C.c_int64) for k in ("tot_mini", "tot_anchors", "tot_chains", "tot_chained", "n_tie_reads", "h2d_bytes", "d2h_bytes")] + \
               [(k, C.c_double) for k in ("sketch_ms", "seed_ms", "sort_ms", "chain_ms")] + [("cells_ref", C.c_int64), ("priv", C.c_void_p)]


class Mm2bError(RuntimeError):
    pass


_lib = None


def load():
    """Load the native library and bind every symbol the header declares; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "16")   # one hardware queue per pipeline stream (see chain_backend.cpp)
    if not os.path.exists(LIB_PATH):
        raise Mm2bError("native library %s not built: run `python __graft_entry__.py` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        getattr(L, name)        # AttributeError if the C ABI is incomplete
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.mm2b_init.restype, L.mm2b_init.argtypes = i32, [i32, vp]
    L.mm2b_init_async.restype, L.mm2b_init_async.argtypes = i32, [i32, vp]
    L.mm2b_shutdown.restype, L.mm2b_shutdown.argtypes = None, []
    L.mm2b_num_devices.restype = i32
    L.mm2b_cuda_device_count.restype = i32
    L.mm2b_last_error.restype = C.c_char_p
    L.mm2b_abi_version.restype = i32
    L.mm2b_host_alloc.restype, L.mm2b_host_alloc.argtypes = vp, [C.c_size_t]
    L.mm2b_host_free.restype, L.mm2b_host_free.argtypes = None, [vp]
    L.mm2b_chain_batch.restype = i32
    L.mm2b_chain_batch.argtypes = [C.POINTER(Params), i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, i64, C.POINTER(Stats)]
    L.mm2b_chain_batch_ex.restype = i32
    L.mm2b_chain_batch_ex.argtypes = [C.POINTER(Params), i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, i64, C.c_uint, C.POINTER(Stats)]
    L.mm2b_chain_batch_device_idx.restype = i32
    L.mm2b_chain_batch_device_idx.argtypes = [vp, C.POINTER(Params), i64, i64] + [vp] * 9 + [vp]
    L.mm2b_unpack_anchors_device.restype = i32
    L.mm2b_unpack_anchors_device.argtypes = [i32, i64, vp, vp, i32, vp, i32, vp, vp]
    L.mm2b_measure_host_copy.restype, L.mm2b_measure_host_copy.argtypes = C.c_double, [i32, C.c_size_t]
    L.mm2b_pack_anchors.restype = i32
    L.mm2b_pack_anchors.argtypes = [vp, i64, vp, vp, C.POINTER(C.c_int32), vp, C.POINTER(C.c_int32), i32]
    L.mm2b_ws_create.restype, L.mm2b_ws_create.argtypes = vp, [i32, i64, i64]
    L.mm2b_ws_destroy.restype, L.mm2b_ws_destroy.argtypes = None, [vp]
    L.mm2b_ws_bytes.restype, L.mm2b_ws_bytes.argtypes = C.c_size_t, [vp]
    L.mm2b_ws_set_counting.restype, L.mm2b_ws_set_counting.argtypes = None, [vp, i32]
    L.mm2b_ws_set_longest_read.restype, L.mm2b_ws_set_longest_read.argtypes = None, [vp, C.c_int64]
    L.mm2b_set_counting.restype, L.mm2b_set_counting.argtypes = None, [i32]
    L.mm2b_chain_batch_device.restype = i32
    L.mm2b_chain_batch_device.argtypes = [vp, C.POINTER(Params), i64, i64] + [vp] * 9 + [vp]
    L.mm2b_ws_stats.restype, L.mm2b_ws_stats.argtypes = i32, [vp, vp, C.POINTER(Stats)]
    L.mm2b_ws_chain_kernel_ms.restype, L.mm2b_ws_chain_kernel_ms.argtypes = C.c_double, [vp]
    L.mm2b_launch_count.restype = i64
    L.mm2b_ws_copy_fpv.restype, L.mm2b_ws_copy_fpv.argtypes = i32, [vp, vp, i64, vp, vp, vp]
    L.mm2b_debug_flags.restype, L.mm2b_debug_flags.argtypes = C.c_uint, []
    L.mm2b_measure_int32_peak.restype, L.mm2b_measure_int32_peak.argtypes = C.c_double, [i32]
    L.mm_chain_dp.restype = vp
    L.mm_chain_dp.argtypes = [i32] * 7 + [C.c_float, i32, i32, i64, vp, C.POINTER(i32), C.POINTER(vp), vp, i32]
    L.mm2b_index_create.restype, L.mm2b_index_create.argtypes = vp, [C.POINTER(IndexDesc)]
    L.mm2b_index_destroy.restype, L.mm2b_index_destroy.argtypes = None, [vp]
    L.mm2b_index_lookup.restype, L.mm2b_index_lookup.argtypes = i32, [vp, i64, vp, vp, vp]
    L.mm2b_map_supported.restype, L.mm2b_map_supported.argtypes = i32, [i32, i32, i32, i32, i64, i32]
    L.mm2b_map_batch.restype = i32
    L.mm2b_map_batch.argtypes = [vp, C.POINTER(SeedParams), C.POINTER(Params), i64, vp, vp, C.POINTER(C.POINTER(MapResult))]
    L.mm2b_map_result_release.restype, L.mm2b_map_result_release.argtypes = None, [C.POINTER(MapResult)]
    L.mm2b_seed_debug.restype = i32
    L.mm2b_seed_debug.argtypes = [vp, C.POINTER(SeedParams), i64, vp, vp] + [C.POINTER(vp)] * 7 + [C.POINTER(i64)]
    L.mm2b_free.restype, L.mm2b_free.argtypes = None, [vp]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise Mm2bError("%s failed (%d): %s" % (what, rc, load().mm2b_last_error().decode()))


def init(devices=None):
    """hardware_init() equivalent. devices: None (all / $MM2B_DEVICES), an int count, or a list of CUDA device ids."""
    L = load()
    if devices is None:
        _check(L.mm2b_init(0, None), "mm2b_init")
    elif isinstance(devices, int):
        _check(L.mm2b_init(devices, None), "mm2b_init")
    else:
        arr = (C.c_int * len(devices))(*devices)
        _check(L.mm2b_init(len(devices), arr), "mm2b_init")


def shutdown():
    load().mm2b_shutdown()


def set_counting(on):
    """Statistics switch for chain_batch()/chain_read(): tally reference-semantics cells (stats.cells_ref); off by default."""
    load().mm2b_set_counting(1 if on else 0)


class PinnedArray:
    """numpy array over pinned host memory from mm2b_host_alloc (H2D/D2H run at PCIe speed only from pinned pages)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = (shape,) if isinstance(shape, int) else tuple(shape)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = load().mm2b_host_alloc(max(n, 1))
        if not self.ptr:
            raise Mm2bError("mm2b_host_alloc(%d) failed: %s" % (n, load().mm2b_last_error().decode()))
        buf = (C.c_char * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            load().mm2b_host_free(self.ptr)
            self.ptr = None


def _p(arr):
    return arr.ctypes.data_as(C.c_void_p)


F_RAW_INPUT, F_DEVICE_GATHER, F_HOST_GATHER = 1, 2, 4


def chain_batch(par, off, a, out=None, want_stats=True, mode="default", flags=0):
    """Chain a CSR batch of reads through the host-buffer batch call (H2D, kernels and D2H inside the call).

    off: int64[n_reads+1]; a: ANCHOR[off[-1]].  `out` may carry preallocated (ideally pinned) arrays
    n_u, n_v, status, u_off, b_off, u, b / bi; otherwise numpy arrays are allocated.  Returns a dict with those plus stats.
    mode: "default" = mm2b_chain_batch (b[] out, the library's default transfer formats);
          "b" / "index" / "both" = mm2b_chain_batch_ex with b[], the int32 indices bi[], or both as outputs and `flags`
          (F_RAW_INPUT, F_DEVICE_GATHER).
    """
    L = load()
    off = np.ascontiguousarray(off, dtype=np.int64)
    a = np.ascontiguousarray(a, dtype=ANCHOR)
    n_reads, n_anchors = len(off) - 1, int(off[-1]) if len(off) else 0
    o = dict(out) if out else {}
    o.setdefault("n_u", np.empty(n_reads, np.int32))
    o.setdefault("n_v", np.empty(n_reads, np.int32))
    o.setdefault("status", np.empty(n_reads, np.int32))
    o.setdefault("u_off", np.empty(n_reads + 1, np.int64))
    o.setdefault("b_off", np.empty(n_reads + 1, np.int64))
    o.setdefault("u", np.empty(max(n_anchors, 1), np.uint64))
    want_b, want_bi = mode in ("default", "b", "both"), mode in ("index", "both")
    if want_b:
        o.setdefault("b", np.empty(max(n_anchors, 1), ANCHOR))
    if want_bi:
        o.setdefault("bi", np.empty(max(n_anchors, 1), np.int32))
    st = Stats()
    if mode == "default":
        rc = L.mm2b_chain_batch(C.byref(par), n_reads, _p(off), _p(a), _p(o["n_u"]), _p(o["n_v"]), _p(o["status"]), _p(o["u_off"]),
                                _p(o["b_off"]), _p(o["u"]), len(o["u"]), _p(o["b"]), len(o["b"]), C.byref(st) if want_stats else None)
    else:
        cap = len(o["b"]) if want_b else len(o["bi"])
        rc = L.mm2b_chain_batch_ex(C.byref(par), n_reads, _p(off), _p(a), _p(o["n_u"]), _p(o["n_v"]), _p(o["status"]), _p(o["u_off"]),
                                   _p(o["b_off"]), _p(o["u"]), len(o["u"]), _p(o["b"]) if want_b else None, _p(o["bi"]) if want_bi else None,
                                   cap, flags, C.byref(st) if want_stats else None)
    _check(rc, "mm2b_chain_batch")
    o["stats"] = st
    return o


def gather_b(off, a, res):
    """b[] from the index output of chain_batch(mode="index"): b[b_off[r]+k] = a[off[r] + bi[b_off[r]+k]] (what a caller that still
    holds a[] does instead of receiving 16-byte copies over PCIe).  Returns an array laid out like res["b"] would be."""
    n_v = res["n_v"].astype(np.int64)
    b = np.zeros(len(res["bi"]), ANCHOR)
    tot = int(n_v.sum())
    if tot:
        rid = np.repeat(np.arange(len(n_v)), n_v)
        k = np.arange(tot) - np.repeat(np.cumsum(n_v) - n_v, n_v)
        pos = res["b_off"][:-1][rid] + k
        b[pos] = a[np.asarray(off[:-1])[rid] + res["bi"][pos]]
    return b


def pack_anchors(a, cap_runs=None):
    """Host side of the packed transfer format (mm2b_pack_anchors): returns (lo[n,2] uint32, xruns[k,2], yruns[m,2])."""
    L = load()
    a = np.ascontiguousarray(a, dtype=ANCHOR)
    n = len(a)
    cap = cap_runs if cap_runs is not None else max(n, 1)
    lo = np.zeros((max(n, 1), 2), np.uint32)
    xr, yr = np.zeros((cap, 2), np.uint32), np.zeros((cap, 2), np.uint32)
    nx, ny = C.c_int32(0), C.c_int32(0)
    _check(L.mm2b_pack_anchors(_p(a), n, _p(lo), _p(xr), C.byref(nx), _p(yr), C.byref(ny), cap), "mm2b_pack_anchors")
    return lo[:n], xr[:nx.value], yr[:ny.value]


def chain_read(par, a):
    """One read through the drop-in mm_chain_dp (km = NULL, so malloc/free like kalloc.c does without an arena).
    Returns (u, b, u_is_null, b_is_null) like the reference's return convention."""
    L = load()
    libc = C.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes = C.c_void_p, [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    a = np.ascontiguousarray(a, dtype=ANCHOR)
    n = len(a)
    pa = None
    if n:
        pa = libc.malloc(n * 16)                   # consumed by mm_chain_dp (chain.c:421)
        C.memmove(pa, a.ctypes.data, n * 16)
    n_u, pu = C.c_int(0), C.c_void_p(0)
    pb = L.mm_chain_dp(par.max_dist_x, par.max_dist_y, par.bw, par.max_skip, par.max_iter, par.min_cnt, par.min_sc,
                       par.gap_scale, par.is_cdna, par.n_segs, n, pa, C.byref(n_u), C.byref(pu), None, 0)
    u = np.empty(n_u.value, np.uint64)
    if n_u.value:
        C.memmove(u.ctypes.data, pu.value, n_u.value * 8)
    n_v = int((u & np.uint64(0xffffffff)).sum())
    b = np.empty(n_v, ANCHOR)
    if n_v:
        C.memmove(b.ctypes.data, pb, n_v * 16)
    u_null, b_null = not pu.value, not pb
    if pu.value:
        libc.free(pu)
    if pb:
        libc.free(pb)
    return u, b, u_null, b_null


class DeviceBatch:
    """A batch whose inputs and outputs live in HBM (torch tensors used purely as device memory + stream handles).

    run() enqueues K0..K3 on the current torch stream through mm2b_chain_batch_device and returns immediately.
    """

    def __init__(self, par, off, a, device=0, keep_fpv=False, index_out=False):
        import torch
        self.torch = torch
        self.L = load()
        self.par = par
        self.device = torch.device("cuda", device)
        off = np.ascontiguousarray(off, dtype=np.int64)
        a = np.ascontiguousarray(a, dtype=ANCHOR)
        self.n_reads, self.n_anchors = len(off) - 1, int(off[-1])
        with torch.cuda.device(self.device):
            self.d_off = torch.from_numpy(off).to(self.device)
            self.d_a = torch.from_numpy(a.view(np.int64).reshape(-1, 2).copy()).to(self.device)
            i32 = dict(dtype=torch.int32, device=self.device)
            i64 = dict(dtype=torch.int64, device=self.device)
            self.d_n_u = torch.empty(max(self.n_reads, 1), **i32)
            self.d_n_v = torch.empty(max(self.n_reads, 1), **i32)
            self.d_status = torch.empty(max(self.n_reads, 1), **i32)
            self.d_u_off = torch.empty(self.n_reads + 1, **i64)
            self.d_b_off = torch.empty(self.n_reads + 1, **i64)
            self.d_u = torch.empty(max(self.n_anchors, 1), **i64)
            self.index_out = index_out
            if index_out:       # chained anchors as int32 indices inside their read (mm2b_chain_batch_device_idx)
                self.d_bi = torch.empty(max(self.n_anchors, 1), **i32)
            else:
                self.d_b = torch.empty((max(self.n_anchors, 1), 2), **i64)
        if keep_fpv:
            os.environ["MM2B_KEEP_FPV"] = "1"
        self.ws = self.L.mm2b_ws_create(device, self.n_anchors, self.n_reads)
        if keep_fpv:
            os.environ.pop("MM2B_KEEP_FPV", None)
        if not self.ws:
            raise Mm2bError("mm2b_ws_create failed: %s" % self.L.mm2b_last_error().decode())

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def run(self):
        fn = self.L.mm2b_chain_batch_device_idx if self.index_out else self.L.mm2b_chain_batch_device
        rc = fn(self.ws, C.byref(self.par), self.n_reads, self.n_anchors,
                self.d_off.data_ptr(), self.d_a.data_ptr(), self.d_n_u.data_ptr(), self.d_n_v.data_ptr(),
                self.d_status.data_ptr(), self.d_u_off.data_ptr(), self.d_b_off.data_ptr(),
                self.d_u.data_ptr(), self.d_bi.data_ptr() if self.index_out else self.d_b.data_ptr(), self._stream())
        _check(rc, "mm2b_chain_batch_device")

    def stats(self):
        st = Stats()
        _check(self.L.mm2b_ws_stats(self.ws, self._stream(), C.byref(st)), "mm2b_ws_stats")
        return st

    def set_counting(self, on):
        self.L.mm2b_ws_set_counting(self.ws, 1 if on else 0)

    def chain_kernel_ms(self):
        return float(self.L.mm2b_ws_chain_kernel_ms(self.ws))

    def fpv(self):
        f, p, v = (np.empty(self.n_anchors, np.int32) for _ in range(3))
        _check(self.L.mm2b_ws_copy_fpv(self.ws, self._stream(), self.n_anchors, _p(f), _p(p), _p(v)), "mm2b_ws_copy_fpv")
        return f, p, v

    def results(self):
        """Copy results to the host (numpy), same keys as chain_batch()."""
        self.torch.cuda.synchronize(self.device)
        n_u = self.d_n_u[:self.n_reads].cpu().numpy()
        n_v = self.d_n_v[:self.n_reads].cpu().numpy()
        u_off, b_off = self.d_u_off.cpu().numpy(), self.d_b_off.cpu().numpy()
        u = self.d_u[:int(u_off[-1])].cpu().numpy().view(np.uint64)
        out = dict(n_u=n_u, n_v=n_v, status=self.d_status[:self.n_reads].cpu().numpy(), u_off=u_off, b_off=b_off, u=u)
        if self.index_out:
            out["bi"] = self.d_bi[:int(b_off[-1])].cpu().numpy()
        else:
            out["b"] = self.d_b[:int(b_off[-1])].cpu().numpy().reshape(-1).view(ANCHOR)
        return out

    def unpack_into_place(self, a):
        """Test hook for the packed transfer format: pack `a` on the host (mm2b_pack_anchors), copy the packed form to the device
        and let mm2b_unpack_anchors_device overwrite this batch's anchors with the restored mm128_t."""
        torch = self.torch
        lo, xr, yr = pack_anchors(a)
        with torch.cuda.device(self.device):
            d_lo = torch.from_numpy(lo.copy().view(np.int32)).to(self.device)
            d_xr = torch.from_numpy(xr.copy().view(np.int32)).to(self.device)
            d_yr = torch.from_numpy(yr.copy().view(np.int32)).to(self.device)
            self.d_a.zero_()
            _check(self.L.mm2b_unpack_anchors_device(self.device.index, len(a), d_lo.data_ptr(), d_xr.data_ptr(), len(xr), d_yr.data_ptr(), len(yr),
                                                     self.d_a.data_ptr(), self._stream()), "mm2b_unpack_anchors_device")
            torch.cuda.synchronize(self.device)
        return self.d_a.cpu().numpy().reshape(-1).view(ANCHOR)[:len(a)]

    def close(self):
        if self.ws:
            self.L.mm2b_ws_destroy(self.ws)
            self.ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- the seeding front end (include/mm2seed_b200.h) ---------------------------------------------------------------------

class Index:
    """The reference's minimizer index on every bound device.  flat: dict(k, w, is_hpc, n_seq, keys, vals, pos) as mm2b_index_flatten
    (host/idx_flatten.cpp) produces it."""

    def __init__(self, flat):
        L = load()
        self.k, self.w = int(flat["k"]), int(flat["w"])
        keys, vals, pos = (np.ascontiguousarray(flat[k], np.uint64) for k in ("keys", "vals", "pos"))
        d = IndexDesc(self.k, self.w, int(flat.get("is_hpc", 0)), int(flat.get("n_seq", 1)), len(keys), len(pos), keys.ctypes.data, vals.ctypes.data, pos.ctypes.data)
        self.h = L.mm2b_index_create(C.byref(d))
        if not self.h:
            raise Mm2bError("mm2b_index_create failed: %s" % L.mm2b_last_error().decode())

    def close(self):
        if self.h:
            load().mm2b_index_destroy(self.h)
            self.h = None

    def lookup(self, minimizers):
        m = np.ascontiguousarray(minimizers, np.uint64)
        n_occ, val = np.zeros(len(m), np.int32), np.zeros(len(m), np.uint64)
        _check(load().mm2b_index_lookup(self.h, len(m), _p(m), _p(n_occ), _p(val)), "mm2b_index_lookup")
        return n_occ, val


def _seq_batch(seqs):
    off = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    return off, b"".join(bytes(s) for s in seqs)


def _take(ptr, dtype, n):
    dt = np.dtype(dtype)
    if n <= 0 or not ptr:
        return np.empty(0, dt)
    return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(ptr), dt, n).copy()


def seed_debug(index, seqs, max_occ):
    """Sketch + seed + sort on the GPU for a small batch; returns the intermediate products (mm2b_seed_debug):
    dict(mini_off, mini, a_off, a, rep_len, n_mini_pos, mini_pos, n_tie_reads)."""
    L = load()
    off, blob = _seq_batch(seqs)
    sp = SeedParams(max_occ, 0)
    ptrs = [C.c_void_p() for _ in range(7)]
    n_tie = C.c_int64(0)
    _check(L.mm2b_seed_debug(index.h, C.byref(sp), len(seqs), _p(off), blob, *[C.byref(p) for p in ptrs], C.byref(n_tie)), "mm2b_seed_debug")
    n = len(seqs)
    mini_off = _take(ptrs[0].value, np.int64, n + 1)
    a_off = _take(ptrs[2].value, np.int64, n + 1)
    out = dict(mini_off=mini_off, mini=_take(ptrs[1].value, ANCHOR, int(mini_off[-1])), a_off=a_off, a=_take(ptrs[3].value, ANCHOR, int(a_off[-1])),
               rep_len=_take(ptrs[4].value, np.int32, n), n_mini_pos=_take(ptrs[5].value, np.int32, n), mini_pos=_take(ptrs[6].value, np.uint32, int(mini_off[-1])),
               n_tie_reads=n_tie.value)
    for p in ptrs:
        L.mm2b_free(p)
    return out


def map_batch(index, seqs, max_occ, par=None, seq_off=None, blob=None, collect=True):
    """mm2b_map_batch: read sequences in, chains out.  seqs: list of bytes (or pass seq_off + blob, a bytes object or a uint8 array).
    Returns a dict of per-read arrays, the call's totals in `stats` and — with collect=True — lists `u`, `b`, `mini_pos` (one numpy
    array per read; for small batches) or — with collect="u" — all u[] entries in read order as one array."""
    L = load()
    if seq_off is None:
        seq_off, blob = _seq_batch(seqs)
    par = par or Params()
    sp = SeedParams(max_occ, 0)
    res = C.POINTER(MapResult)()
    blob_p = blob if isinstance(blob, (bytes, bytearray)) or blob is None else blob.ctypes.data_as(C.c_void_p)
    _check(L.mm2b_map_batch(index.h, C.byref(sp), C.byref(par), len(seq_off) - 1, _p(seq_off), blob_p, C.byref(res)), "mm2b_map_batch")
    r = res.contents
    n = int(r.n_reads)
    out = {k: np.ctypeslib.as_array(getattr(r, k), (n,)).copy() if n else np.empty(0, np.int64) for k in ("status", "n_u", "n_v", "rep_len", "n_mini_pos", "n_mini", "seg", "n_a", "u_off", "b_off", "mp_off")}
    out["stats"] = {k: getattr(r, k) for k in ("tot_mini", "tot_anchors", "tot_chains", "tot_chained", "n_tie_reads", "h2d_bytes", "d2h_bytes", "sketch_ms", "seed_ms", "sort_ms", "chain_ms", "cells_ref")}
    out["stats"]["n_segs"] = int(r.n_segs)
    if collect == "u":                                  # every u[] entry in read order (segment by segment, vectorised)
        parts = []
        n_u, seg, u_off = out["n_u"].astype(np.int64), out["seg"], out["u_off"]
        for sg in range(int(r.n_segs)):
            sel = np.flatnonzero(seg == sg)
            cnt = n_u[sel]
            tot = int(cnt.sum())
            if not tot:
                continue
            hi = int((u_off[sel] + cnt).max())
            arr = _take(r.seg_u[sg], np.uint64, hi)
            idx = np.repeat(u_off[sel] - (np.cumsum(cnt) - cnt), cnt) + np.arange(tot)
            parts.append(arr[idx])
        out["u_flat"] = np.concatenate(parts) if parts else np.empty(0, np.uint64)
        collect = False
    if not collect:
        L.mm2b_map_result_release(res)
        return out
    u, b, mp = [], [], []
    for i in range(n):
        s = int(out["seg"][i])
        u.append(_take((r.seg_u[s] or 0) + 8 * int(out["u_off"][i]), np.uint64, int(out["n_u"][i])))
        b.append(_take((r.seg_b[s] or 0) + 16 * int(out["b_off"][i]), ANCHOR, int(out["n_v"][i])))
        mp.append(_take((r.seg_mini_pos[s] or 0) + 4 * int(out["mp_off"][i]), np.uint32, int(out["n_mini_pos"][i])))
    out["u"], out["b"], out["mini_pos"] = u, b, mp
    L.mm2b_map_result_release(res)
    return out
