"""TEST INFRASTRUCTURE — reader for captured mm_chain_dp calls (format: oracle/dump_format.h)."""
import gzip
import struct

import numpy as np

from .oracle_py import ANCHOR, Params

MAGIC = 0x4443324D
HAS_FPV, B_NULL, U_NULL = 1, 2, 4
_HDR = struct.Struct("<II9ifqii")
assert _HDR.size == 64


def read_dump(path):
    """Return a list of records: dict(par, a, f, p, v, u, b, b_null, u_null)."""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as fh:
        buf = fh.read()
    out, pos = [], 0
    while pos < len(buf):
        (magic, flags, mdx, mdy, bw, skip, it, cnt, sc, cdna, segs, gs, n, n_u, n_v) = _HDR.unpack_from(buf, pos)
        assert magic == MAGIC, "bad dump record at %d" % pos
        pos += 64
        a = np.frombuffer(buf, ANCHOR, n, pos); pos += 16 * n
        f = p = v = None
        if flags & HAS_FPV:
            fpv = np.frombuffer(buf, "<i4", 3 * n, pos); pos += 12 * n
            f, p, v = fpv[:n], fpv[n:2 * n], fpv[2 * n:]
        u = np.frombuffer(buf, "<u8", n_u, pos); pos += 8 * n_u
        b = np.frombuffer(buf, ANCHOR, n_v, pos); pos += 16 * n_v
        out.append(dict(par=Params(mdx, mdy, bw, skip, it, cnt, sc, cdna, segs, gs), a=a, f=f, p=p, v=v, u=u, b=b,
                        b_null=bool(flags & B_NULL), u_null=bool(flags & U_NULL)))
    return out


def to_batch(records):
    """Pack records that share parameters into a CSR batch (off[n_reads+1], a[total])."""
    off = np.zeros(len(records) + 1, np.int64)
    np.cumsum([len(r["a"]) for r in records], out=off[1:])
    a = np.concatenate([r["a"] for r in records]) if records else np.empty(0, ANCHOR)
    return off, a
