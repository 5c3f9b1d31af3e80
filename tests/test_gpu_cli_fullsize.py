"""BASELINE.json configs[1] / configs[2] at size, through the drop-in itself: the reference CLI linked against the B200 backend
(oracle/_ref/minimap2-b200, chain.c replaced by libmm2chain_b200's mm_chain_dp) must print byte-identical PAF to the reference CLI
with its own software chaining (oracle/_ref/minimap2-sw) for

  * 50,000 synthetic CCS-like reads (15 kb, ~1 % error), -x asm20      (configs[2]: "bit-exact PAF check")
  * 20,000 synthetic ONT reads (10 kb, ~10 % error), -x map-ont        (configs[1] shape)

against the 100 Mbp random reference — both through the per-read drop-in (minimap2-b200: mm_chain_dp on the GPU behind the cross-thread
batcher) and through the phase-split caller with the seeding front end (minimap2-b200-batch: sketch, seed hits, sort and chaining of a
whole mini-batch on the GPU, include/mm2seed_b200.h).  The binaries are built in place from /root/reference by oracle/Makefile and travel with
the repository; the inputs are simulated here (bench_workloads.py), nothing reads /root/reference at run time.
MM2B_TEST_CCS_READS / MM2B_TEST_ONT_READS shrink the cases for quick runs.
"""
import hashlib
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
SW = os.path.join(ROOT, "oracle", "_ref", "minimap2-sw")
B200 = os.path.join(ROOT, "oracle", "_ref", "minimap2-b200")
BATCH = os.path.join(ROOT, "oracle", "_ref", "minimap2-b200-batch")


def _run(exe, args, threads, env=None):
    p = subprocess.run([exe, "-t", str(threads)] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1500, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return hashlib.md5(p.stdout).hexdigest(), p.stdout.count(b"\n")


@pytest.mark.parametrize("name,n_reads_env,default_reads", [("asm20", "MM2B_TEST_CCS_READS", 50000), ("map-ont", "MM2B_TEST_ONT_READS", 20000)])
def test_cli_paf_identical_at_size(pkg, tmp_path, name, n_reads_env, default_reads):
    if not (os.path.exists(SW) and os.path.exists(B200)):
        pytest.skip("oracle/_ref CLIs were not built (needs /root/reference at build time)")
    sys.path.insert(0, ROOT)
    import bench_workloads as BW
    n_reads = int(os.environ.get(n_reads_env, default_reads))
    threads = os.cpu_count() or 8
    fa, mmi = BW._reference_files(name, pkg("seqsim"), threads)
    q = str(tmp_path / "reads.fa")
    BW._simulate_reads(name, n_reads, 77, q, procs=min(threads, 16))
    args = BW.PRESETS[name][2] + [mmi, q]
    md5_sw, lines_sw = _run(SW, args, threads)
    assert lines_sw >= 0.95 * n_reads                                    # the simulated reads do map
    # an oversubscribed -t keeps hundreds of reads in flight at the per-read boundary (the cross-thread batcher aggregates them)
    md5_gpu, lines_gpu = _run(B200, args, 256)
    assert (md5_gpu, lines_gpu) == (md5_sw, lines_sw), "%s: PAF of the B200 drop-in differs from the reference's software chaining" % name
    if os.path.exists(BATCH):
        md5_b, lines_b = _run(BATCH, args, threads)
        assert (md5_b, lines_b) == (md5_sw, lines_sw), "%s: PAF of the phase-split caller with the GPU seeding front end differs from the reference's" % name
        md5_b, lines_b = _run(BATCH, args, threads, env={"MM2B_FRONT": "0"})
        assert (md5_b, lines_b) == (md5_sw, lines_sw), "%s: PAF of the phase-split caller (host seeding, GPU chaining) differs from the reference's" % name
