"""The fiber-based kt_for() (minimap2-fpga_b200/host/fiber_for.cpp) under the reference CLI.

CPU: oracle/_ref/minimap2-fiber-sw = the reference CLI with kthread.c's kt_for renamed away (a build flag), the product's
fiber scheduler in its place and the reference's own chain.c behind the park / resume protocol (oracle/fiber_sw_shim.cpp).  Its
PAF must be byte-identical to the reference's for every captured case at 1, 4 and 64 "threads" (= fibers), which checks the
scheduler, the tid / kalloc-arena discipline and the second chaining call of the short-read preset on real mapping runs.
GPU: oracle/_ref/minimap2-fiber-b200, the same with the B200 backend chaining each OS thread's parked reads in one batch.
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

SW = os.path.join(ROOT, "oracle", "_ref", "minimap2-fiber-sw")
B200 = os.path.join(ROOT, "oracle", "_ref", "minimap2-fiber-b200")


@pytest.fixture(scope="module")
def cases():
    sys.path.insert(0, GOLDEN)
    import cases as golden_cases
    with tempfile.TemporaryDirectory() as td:
        yield golden_cases.build_cases(td)


def _paf(exe, args, threads, env=None):
    out = subprocess.run([exe, "-t", str(threads)] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout


def _check(exe, cases, threads, only=None, env=None):
    gold = json.load(open(os.path.join(GOLDEN, "paf_md5.json")))
    bad = []
    for name, args in cases:
        if only and name not in only:
            continue
        paf = _paf(exe, args, threads, env)
        if hashlib.md5(paf).hexdigest() != gold[name]["md5"] or paf.count(b"\n") != gold[name]["lines"]:
            bad.append(name)
    assert not bad, "PAF differs from the reference at -t %d for: %s" % (threads, bad)


def test_scheduler_unit(tmp_path):
    """tests/native/fiber_unit.cpp: every item once, one item per tid at a time, results reach the right fiber; sync and async."""
    exe = str(tmp_path / "fiber_unit")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "minimap2-fpga_b200", "host"),
                           os.path.join(ROOT, "tests", "native", "fiber_unit.cpp"), os.path.join(ROOT, "minimap2-fpga_b200", "host", "fiber_for.cpp"),
                           "-o", exe, "-lpthread"])
    for env in ({}, {"MM2B_FIBER_OS_THREADS": "1"}, {"MM2B_FIBER_OS_THREADS": "5", "MM2B_FIBER_STACK_KB": "64"}):
        out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **env), timeout=300)
        assert out.returncode == 0 and b"FAILED" not in out.stdout and out.stdout.count(b": ok") == 48, (env, out.stdout.decode()[-1500:], out.stderr.decode()[-500:])


@pytest.mark.parametrize("threads", [1, 4, 64])
def test_reference_cli_on_fibers_with_software_chaining(cases, threads):
    if not os.path.exists(SW):
        pytest.skip("oracle/_ref/minimap2-fiber-sw was not built (needs /root/reference at build time)")
    _check(SW, cases, threads)


def test_more_fibers_than_reads_and_two_os_threads(cases):
    if not os.path.exists(SW):
        pytest.skip("oracle/_ref/minimap2-fiber-sw was not built (needs /root/reference at build time)")
    _check(SW, cases, 300, only=("mt_map-ont", "inv_map-ont", "sr_paired", "syn_ont"), env={"MM2B_FIBER_OS_THREADS": "2"})
    _check(SW, cases, 7, only=("syn_ccs", "splice", "tandem_iter64"), env={"MM2B_FIBER_OS_THREADS": "3", "MM2B_FIBER_STACK_KB": "256"})


@pytest.mark.parametrize("threads", [2, 9, 64])
def test_two_groups_of_fibers_against_an_asynchronous_backend(cases, threads):
    """submit() / wait(): a batch goes out when half of an OS thread's fibers are parked, the other half keeps running."""
    if not os.path.exists(SW):
        pytest.skip("oracle/_ref/minimap2-fiber-sw was not built (needs /root/reference at build time)")
    _check(SW, cases, threads, env={"MM2_FIBER_SHIM_ASYNC": "1"})
    _check(SW, cases, threads, only=("syn_ont", "sr_paired", "tandem_iter64"), env={"MM2_FIBER_SHIM_ASYNC": "1", "MM2B_FIBER_OS_THREADS": "1"})


@pytest.mark.gpu
def test_reference_cli_on_fibers_with_the_b200_backend(cases):
    if not os.path.exists(B200):
        pytest.skip("oracle/_ref/minimap2-fiber-b200 was not built (needs /root/reference at build time)")
    _check(B200, cases, 256)
    _check(B200, cases, 8, only=("syn_ont", "sr_paired", "tandem_iter64"))
