"""GPU test of the drop-in itself: the reference minimap2 CLI (main.c, map.c, format.c ... compiled unchanged from
/root/reference into oracle/_ref/minimap2-b200 by oracle/Makefile) with chain.c REPLACED by libmm2chain_b200's mm_chain_dp.
Its PAF output must be byte-identical to what the reference's software chaining printed for the same command
(md5s in tests/golden/paf_md5.json, written by tests/golden/make_golden.py).  Nothing here reads /root/reference.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "oracle", "_ref", "minimap2-b200")


@pytest.fixture(scope="module")
def cases():
    if not os.path.exists(CLI):
        pytest.skip("oracle/_ref/minimap2-b200 was not built (needs /root/reference at build time)")
    sys.path.insert(0, GOLDEN)
    import cases as golden_cases
    with tempfile.TemporaryDirectory() as td:
        yield golden_cases.build_cases(td)


def _paf(args, threads):
    out = subprocess.run([CLI, "-t", str(threads)] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout


def test_paf_is_byte_identical_to_reference_software_chaining(cases):
    gold = json.load(open(os.path.join(GOLDEN, "paf_md5.json")))
    bad = []
    for name, args in cases:
        paf = _paf(args, threads=1)
        if hashlib.md5(paf).hexdigest() != gold[name]["md5"] or paf.count(b"\n") != gold[name]["lines"]:
            bad.append(name)
    assert not bad, "PAF differs from the reference for: %s" % bad


def test_paf_is_thread_count_independent(cases):
    """kt_for worker threads (map.c:561) call mm_chain_dp concurrently; each gets its own stream and workspace."""
    gold = json.load(open(os.path.join(GOLDEN, "paf_md5.json")))
    for name, args in cases:
        if name in ("syn_ont", "syn_ccs", "sr_paired", "inv_map-ont", "tandem_iter64"):
            paf = _paf(args, threads=8)
            assert hashlib.md5(paf).hexdigest() == gold[name]["md5"], name
