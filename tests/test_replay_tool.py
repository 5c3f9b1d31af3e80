"""mm2b-replay (minimap2-fpga_b200/host/replay_main.cpp), the batched caller for anchor dumps.

CPU: the binary is built, its dump header (include/mm2chain_dump.h) is the format the golden fixtures are written in
(oracle/dump_format.h, written by oracle/dump_shim.c around the REFERENCE's mm_chain_dp), bad input is refused and a machine
without a GPU is an error (no CPU fallback).  GPU: every golden dump replays to exactly the recorded reference results.
"""
import gzip
import os
import struct
import subprocess
import tempfile

import pytest

from conftest import GOLDEN, ROOT, golden_names

TOOL = os.path.join(ROOT, "minimap2-fpga_b200", "mm2b-replay")


@pytest.fixture(scope="module")
def tool(pkg):
    pkg("build").build_all()
    assert os.path.exists(TOOL)
    return TOOL


def _run(args, **kw):
    return subprocess.run([TOOL] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, **kw)


def test_dump_header_is_the_fixture_format(tmp_path):
    src = tmp_path / "fmt.c"
    src.write_text('''
#include <stddef.h>
#include "mm2chain_dump.h"
#include "dump_format.h"
#define SAME(f) _Static_assert(offsetof(mm2b_dump_hdr_t, f) == offsetof(mm2_dump_hdr_t, f), #f)
_Static_assert(sizeof(mm2b_dump_hdr_t) == 64 && sizeof(mm2_dump_hdr_t) == 64, "size");
SAME(magic); SAME(flags); SAME(max_dist_x); SAME(max_dist_y); SAME(bw); SAME(max_skip); SAME(max_iter); SAME(min_cnt); SAME(min_sc);
SAME(is_cdna); SAME(n_segs); SAME(gap_scale); SAME(n); SAME(n_u); SAME(n_v);
_Static_assert(MM2B_DUMP_MAGIC == MM2_DUMP_MAGIC && MM2B_DUMP_HAS_FPV == MM2_DUMP_HAS_FPV && MM2B_DUMP_B_NULL == MM2_DUMP_B_NULL
               && MM2B_DUMP_U_NULL == MM2_DUMP_U_NULL, "constants");
int main(void) { return 0; }
''')
    subprocess.check_call(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"), "-c", str(src),
                           "-o", str(tmp_path / "fmt.o")])


def test_usage_and_bad_dumps_are_refused(tool, tmp_path):
    assert _run([]).returncode == 2
    assert _run(["--no-such-flag", "x"]).returncode == 2
    assert _run([str(tmp_path / "missing.dump")]).returncode == 2
    bad = tmp_path / "bad.dump"
    bad.write_bytes(b"\0" * 64)
    r = _run([str(bad)])
    assert r.returncode == 2 and b"bad record header" in r.stderr
    raw = gzip.open(os.path.join(GOLDEN, "mt_map-ont.dump.gz")).read()
    cut = tmp_path / "cut.dump"
    cut.write_bytes(raw[:200])
    r = _run([str(cut)])
    assert r.returncode == 2 and b"truncated" in r.stderr


def test_no_gpu_is_an_error_not_a_fallback(tool):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = _run(["--check", os.path.join(GOLDEN, "mt_map-ont.dump.gz")], env=env)
    assert r.returncode == 3 and b"no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_every_golden_dump_replays_to_the_recorded_reference_results(tool):
    paths = [os.path.join(GOLDEN, n + ".dump.gz") for n in golden_names()]
    r = _run(["--check", "-r", "2"] + paths)
    assert r.returncode == 0, (r.stdout.decode()[-3000:], r.stderr.decode()[-2000:])
    out = r.stdout.decode()
    assert "MISMATCH" not in out and "mismatching reads 0" in out
    # the same through many small calls (sub-batches of 7 reads)
    r = _run(["--check", "--quiet", "-B", "7"] + paths[:6])
    assert r.returncode == 0 and "mismatching reads 0" in r.stdout.decode()


@pytest.mark.gpu
def test_a_tampered_dump_is_reported(tool, tmp_path):
    raw = bytearray(gzip.open(os.path.join(GOLDEN, "mt_map-ont.dump.gz")).read())
    magic, flags, *_rest = struct.unpack_from("<II9ifqii", raw, 0)
    n, n_u, n_v = struct.unpack_from("<qii", raw, 48)
    assert n_u > 0
    pos_u = 64 + 16 * n + (12 * n if flags & 1 else 0)
    raw[pos_u + 4] ^= 1                      # one bit of the first chain's score
    p = tmp_path / "tampered.dump"
    p.write_bytes(bytes(raw))
    r = _run(["--check", str(p)])
    assert r.returncode == 1 and "MISMATCH" in r.stdout.decode()
