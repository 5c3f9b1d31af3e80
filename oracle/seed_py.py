"""TEST INFRASTRUCTURE — python access to the seeding oracle (oracle/seed_oracle.c) and to the files written by the reference-side
tool oracle/_ref/mm2-seed-ref (oracle/seed_ref_tool.cpp): what the reference's own mm_sketch / collect_seed_hits produce per read,
and the flat index.  Only tests/, smoke() and bench.py's CPU legs import this."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

from . import oracle_py as O

HERE = os.path.dirname(os.path.abspath(__file__))
TOOL = os.path.join(HERE, "_ref", "mm2-seed-ref")
ANCHOR = O.ANCHOR


def have_tool():
    return os.path.exists(TOOL)


def run_tool(preset, ref, reads, seeds_out, index_out="-", mid_occ=None):
    cmd = [TOOL, preset, ref, reads, seeds_out, index_out] + ([str(mid_occ)] if mid_occ is not None else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


def read_index(path):
    """-> dict(k, w, is_hpc, n_seq, mid_occ, keys, vals, pos)"""
    with open(path, "rb") as fh:
        magic, k, w, hpc, n_seq, mid_occ = struct.unpack("<6i", fh.read(24))
        assert magic == 0x5849324D, "not a flat index file"
        n_keys, n_pos = struct.unpack("<2q", fh.read(16))
        keys = np.fromfile(fh, "<u8", n_keys)
        vals = np.fromfile(fh, "<u8", n_keys)
        pos = np.fromfile(fh, "<u8", n_pos)
    return dict(k=k, w=w, is_hpc=hpc, n_seq=n_seq, mid_occ=mid_occ, keys=keys, vals=vals, pos=pos)


def read_seeds(path):
    """-> dict(k, w, mid_occ, reads=[dict(qlen, mv, a, rep_len, mini_pos)])"""
    buf = open(path, "rb").read()
    magic, k, w, mid_occ = struct.unpack_from("<4i", buf, 0)
    assert magic == 0x5332324D, "not a seed record file"
    p, reads = 16, []
    while p < len(buf):
        qlen, n_mv, rep_len, n_mp = struct.unpack_from("<4i", buf, p)
        n_a, = struct.unpack_from("<q", buf, p + 16)
        p += 24
        mv = np.frombuffer(buf, ANCHOR, n_mv, p); p += 16 * n_mv
        a = np.frombuffer(buf, ANCHOR, n_a, p); p += 16 * n_a
        mp = np.frombuffer(buf, "<u8", n_mp, p); p += 8 * n_mp
        reads.append(dict(qlen=qlen, mv=mv, a=a, rep_len=rep_len, mini_pos=mp))
    return dict(k=k, w=w, mid_occ=mid_occ, reads=reads)


_sig = False


def _lib():
    global _sig
    L = O.lib()
    if not _sig:
        L.mm2o_sketch.restype = C.c_int64
        L.mm2o_sketch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.mm2o_index_new.restype = C.c_void_p
        L.mm2o_index_new.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mm2o_index_free.argtypes = [C.c_void_p]
        L.mm2o_index_get.restype = C.c_int
        L.mm2o_index_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.mm2o_seed.restype = C.c_int64
        L.mm2o_seed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]
        _sig = True
    return L


def sketch(seq, w, k):
    """mm_sketch restated (non-HPC): seq bytes -> ANCHOR array of minimizers."""
    L = _lib()
    out = np.empty(len(seq) + 8, ANCHOR)
    n = L.mm2o_sketch(bytes(seq), len(seq), w, k, out.ctypes.data, len(out))
    assert n >= 0
    return out[:n].copy()


class Index:
    def __init__(self, flat):
        self.flat = flat          # keeps the arrays alive
        self.h = _lib().mm2o_index_new(flat["k"], flat["w"], len(flat["keys"]), flat["keys"].ctypes.data, flat["vals"].ctypes.data, flat["pos"].ctypes.data)

    def __del__(self):
        try:
            _lib().mm2o_index_free(self.h)
        except Exception:
            pass

    def get(self, minimizer):
        v = C.c_uint64(0)
        n = _lib().mm2o_index_get(self.h, int(minimizer), C.byref(v))
        return n, v.value

    def seed(self, mv, qlen, max_occ, cap=None):
        """collect_matches + collect_seed_hits restated -> (a, rep_len, mini_pos)"""
        L = _lib()
        mv = np.ascontiguousarray(mv, ANCHOR)
        cap = cap or max(1024, 64 * len(mv))
        while True:
            a = np.empty(cap, ANCHOR)
            mp = np.empty(max(len(mv), 1), np.uint64)
            rep, nmp = C.c_int32(0), C.c_int32(0)
            n = L.mm2o_seed(self.h, max_occ, qlen, len(mv), mv.ctypes.data, a.ctypes.data, cap, C.byref(rep), C.byref(nmp), mp.ctypes.data)
            if n >= 0:
                return a[:n].copy(), rep.value, mp[:nmp.value].copy()
            cap *= 8
