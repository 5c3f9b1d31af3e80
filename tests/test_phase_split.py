"""The phase-split caller (minimap2-fpga_b200/host/map_batch.cpp) under the reference CLI.

CPU: oracle/_ref/minimap2-batch-sw = the reference CLI with map.c compiled through map_batch.cpp (seed all reads of a mini-batch,
chain them in batch calls, finish all reads) over a software stand-in for the batch call (oracle/batch_sw_shim.cpp: the reference's
own chain.c per staged read).  Its PAF must be byte-identical to the reference's for every captured case — this checks the staging
blocks, the CSR layout, the index gather, the kalloc discipline of the two halves of mm_map_frag, mate flipping of paired reads
and the second chaining pass of the short-read preset on real mapping runs, without a GPU.
GPU: oracle/_ref/minimap2-b200-batch, the same caller over libmm2chain_b200.so.
Nothing here reads /root/reference (the binaries are prebuilt by oracle/Makefile).
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import pytest

from conftest import GOLDEN, ROOT

SW = os.path.join(ROOT, "oracle", "_ref", "minimap2-batch-sw")
B200 = os.path.join(ROOT, "oracle", "_ref", "minimap2-b200-batch")


@pytest.fixture(scope="module")
def cases():
    sys.path.insert(0, GOLDEN)
    import cases as golden_cases
    with tempfile.TemporaryDirectory() as td:
        yield golden_cases.build_cases(td)


def _paf(exe, args, threads, extra=(), env=None):
    out = subprocess.run([exe, "-t", str(threads)] + list(extra) + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout


def _check(exe, cases, threads, only=None, extra=(), env=None):
    gold = json.load(open(os.path.join(GOLDEN, "paf_md5.json")))
    bad = []
    for name, args in cases:
        if only and name not in only:
            continue
        paf = _paf(exe, args, threads, extra, env)
        if hashlib.md5(paf).hexdigest() != gold[name]["md5"] or paf.count(b"\n") != gold[name]["lines"]:
            bad.append(name)
    assert not bad, "PAF differs from the reference at -t %d for: %s" % (threads, bad)


@pytest.mark.parametrize("threads", [1, 5])
def test_phase_split_with_software_chaining(cases, threads):
    if not os.path.exists(SW):
        pytest.skip("oracle/_ref/minimap2-batch-sw was not built (needs /root/reference at build time)")
    _check(SW, cases, threads)


def test_phase_split_many_small_mini_batches(cases):
    """-K 1k: a mini-batch per read or two, so staging blocks are recycled hundreds of times."""
    if not os.path.exists(SW):
        pytest.skip("oracle/_ref/minimap2-batch-sw was not built (needs /root/reference at build time)")
    _check(SW, cases, 3, only=("syn_ont", "sr_paired", "splice", "inv_map-ont", "ava"), extra=("-K", "1k"))


def test_phase_split_switched_off_is_the_reference_worker(cases):
    if not os.path.exists(SW):
        pytest.skip("oracle/_ref/minimap2-batch-sw was not built (needs /root/reference at build time)")
    _check(SW, cases, 2, only=("syn_ont", "sr_paired"), env={"MM2B_PHASE_SPLIT": "0"})


def _front_end_lines(exe, args):
    out = subprocess.run([exe, "-t", "3"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900, env=dict(os.environ, MM2B_TRACE="1"))
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return [l for l in out.stderr.decode().splitlines() if "front end:" in l]


def test_seeding_front_end_is_taken_where_it_applies_and_only_there(cases):
    """map-ont / asm20 go through mm2b_map_batch (sequences in, chains out); paired short reads, splice-free presets with per-read
    chaining gaps and all-vs-all overlap (seed-skipping flags) keep the host seeding.  MM2B_FRONT=0 switches it off: same PAF."""
    if not os.path.exists(SW):
        pytest.skip("oracle/_ref/minimap2-batch-sw was not built (needs /root/reference at build time)")
    by_name = dict(cases)
    assert _front_end_lines(SW, by_name["syn_ont"]) and _front_end_lines(SW, by_name["syn_ccs"]) and _front_end_lines(SW, by_name["tandem_iter64"])
    assert not _front_end_lines(SW, by_name["sr_paired"]) and not _front_end_lines(SW, by_name["ava"])
    _check(SW, cases, 4, only=("syn_ont", "syn_ccs", "inv_map-ont", "tandem_iter64"), env={"MM2B_FRONT": "0"})


@pytest.mark.gpu
def test_seeding_front_end_on_the_b200(cases):
    if not os.path.exists(B200):
        pytest.skip("oracle/_ref/minimap2-b200-batch was not built (needs /root/reference at build time)")
    by_name = dict(cases)
    assert _front_end_lines(B200, by_name["syn_ont"]) and _front_end_lines(B200, by_name["tandem_iter64"])
    _check(B200, cases, 8, only=("syn_ont", "syn_ccs"), env={"MM2B_FRONT": "0"})


@pytest.mark.gpu
def test_front_end_many_small_mini_batches_on_the_b200(cases):
    """-K 1k / -K 20k: hundreds of tiny mm2b_map_batch calls, so contexts, pinned segments and the staging buffer are recycled."""
    if not os.path.exists(B200):
        pytest.skip("oracle/_ref/minimap2-b200-batch was not built (needs /root/reference at build time)")
    _check(B200, cases, 4, only=("syn_ont", "syn_ccs", "inv_map-ont", "tandem"), extra=("-K", "1k"))
    _check(B200, cases, 4, only=("syn_ont", "splice"), extra=("-K", "20k"))


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 8])
def test_phase_split_with_the_b200_backend(cases, threads):
    if not os.path.exists(B200):
        pytest.skip("oracle/_ref/minimap2-b200-batch was not built (needs /root/reference at build time)")
    _check(B200, cases, threads)
