"""CPU test of the N>1 host logic (world_size 2, gloo): shard plan, per-rank chaining, ordered gather on rank 0.
The per-shard chaining function here is the oracle (the checker); on GPUs bench/production pass binding.chain_batch."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import fuzz
from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, seed, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from __graft_entry__ import load_package
    from oracle import oracle_py as O
    sharding = load_package("sharding")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, a = fuzz.mixed_batch(seed, n_reads=40, scale=0.4)

    def chain_fn(off_s, a_s):
        r = O.replay(O.Params(), off_s, a_s, n_threads=2)
        return dict(n_u=r["n_u"], n_v=r["n_v"], u_off=off_s[:-1], b_off=off_s[:-1], u=r["u"], b=r["b"])

    out = sharding.run_sharded(chain_fn, off, a, dist)
    if rank == 0:
        q.put([(u.tolist(), b.tobytes()) for u, b in zip(*out)])
    dist.barrier()
    dist.destroy_process_group()


def test_plan_is_contiguous_balanced_and_total(pkg):
    sharding = pkg("sharding")
    off, a = fuzz.mixed_batch(3, n_reads=50, scale=0.3)
    for world in (1, 2, 3, 8, 64):
        b = sharding.plan(off, world)
        assert b[0] == 0 and b[-1] == len(off) - 1 and np.all(np.diff(b) >= 0) and len(b) == world + 1
        got = [sharding.shard(off, a, world, r) for r in range(world)]
        assert sum(len(g[1]) for g in got) == len(a)
        assert np.array_equal(np.concatenate([g[1] for g in got]), a)
        sizes = [len(g[1]) for g in got]
        if world <= 8:
            assert max(sizes) - min(sizes) <= 2 * int(np.diff(off).max())
    off0 = np.zeros(1, np.int64)
    assert list(sharding.plan(off0, 4)) == [0, 0, 0, 0, 0]


def test_world2_gloo_gather_equals_unsharded(oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 17, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    off, a = fuzz.mixed_batch(17, n_reads=40, scale=0.4)
    ref = oracle.replay(oracle.Params(), off, a, n_threads=2)
    assert len(got) == len(off) - 1
    for r, (u, b) in enumerate(got):
        o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
        assert u == ref["u"][o:o + nu].tolist() and b == ref["b"][o:o + nv].tobytes(), r
