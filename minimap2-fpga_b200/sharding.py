"""Read sharding for multi-process (one rank per GPU) runs: reads are independent units of mm_chain_dp, so a batch is cut
into `world` contiguous read ranges balanced by anchor count; every rank chains its own range and the per-read results
are gathered in input order on rank 0.  No data-path collective: the only communication is the final gather of results
(torch.distributed as plumbing; NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def plan(off, world):
    """Boundaries r[0..world] (read indices) of `world` contiguous shards with near-equal anchor counts."""
    off = np.asarray(off, dtype=np.int64)
    n_reads, total = len(off) - 1, int(off[-1])
    targets = (np.arange(1, world, dtype=np.float64) * total / world)
    cuts = np.searchsorted(off, targets, side="left")
    bounds = np.concatenate([[0], np.clip(cuts, 0, n_reads), [n_reads]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def shard(off, a, world, rank):
    """Rank's CSR slice (offsets rebased to 0) and its read range."""
    b = plan(off, world)
    r0, r1 = int(b[rank]), int(b[rank + 1])
    a0, a1 = int(off[r0]), int(off[r1])
    return off[r0:r1 + 1] - a0, a[a0:a1], (r0, r1)


def run_sharded(chain_fn, off, a, dist=None):
    """chain_fn(off_slice, a_slice) -> dict(n_u, n_v, u_off, b_off, u, b) for the slice.  Returns on rank 0 the per-read
    lists (u_list, b_list) for the whole batch in input order, None elsewhere."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    off_s, a_s, (r0, r1) = shard(off, a, world, rank)
    res = chain_fn(off_s, a_s)
    mine = []
    for r in range(r1 - r0):
        nu, nv, uo, bo = int(res["n_u"][r]), int(res["n_v"][r]), int(res["u_off"][r]), int(res["b_off"][r])
        mine.append((np.array(res["u"][uo:uo + nu]), np.array(res["b"][bo:bo + nv])))
    if world == 1:
        return [m[0] for m in mine], [m[1] for m in mine]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    flat = [x for part in gathered for x in part]        # rank order == read order: shards are contiguous
    return [m[0] for m in flat], [m[1] for m in flat]
