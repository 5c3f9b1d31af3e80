#!/usr/bin/env python
"""Benchmark of the chaining hot path (mm_chain_dp) on B200 — contract in the task brief.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl b200|reference] [--workload map-ont|asm20|ultralong|tandem]

One "step" = one pass of the hot path over one batch of reads (BASELINE.json configs[1]: map-ont, 100k synthetic ONT reads of
10 kb mean / ~10 % error vs a 100 Mbp random reference, per GPU: weak scaling).  The batch holds the anchors the REFERENCE's own
seeding hands to mm_chain_dp for those reads (bench_workloads.py: the reference CLI built in place seeds the simulated reads once,
untimed, and the capture shim records every call; `--source model` falls back to the calibrated anchor model of round 1).
  value   GCUPS with anchors already resident in HBM (order + chaining kernels on the device, CUDA-event timed)
  e2e     GCUPS through the host-buffer C-ABI call mm2b_chain_batch: pinned host anchors (mm128_t) in, u[] and b[] (mm128_t)
          out, host<->device copies and the library's host-side packing / gathering inside the timed region; the index-output
          variant (mm2b_chain_batch_ex, bi[] instead of b[]) and the round-1 transfer format (raw 16 B in, b[] copied back from
          the device) are timed beside it in e2e.variants.  At N > 1 `e2e` is ONE mm2b_chain_batch call made by rank 0 over all
          N devices (the library's own per-device worker threads), on N x reads; the per-process number is kept beside it.
GCUPS counts reference-semantics cells (iterations of chain.c:197), which the kernel tallies exactly (tests check the tally
against the oracle).  The reference arm (--impl reference) and the cpu_baseline object time the reference's own compiled
software chaining (oracle/_ref/libmm2ref.so, else the oracle port) on all host cores over the same reads.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "16")   # before any CUDA context exists: one hardware queue per pipeline stream
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
import bench_workloads as BW  # noqa: E402

INT_OPS_PER_CELL = 30          # INT32-equivalent ops per reference cell
BYTES_PER_ANCHOR = 40          # unavoidable device traffic per anchor (16 in, <=16 b out, <=8 u/indices)
KERNEL_SRC = os.path.join(ROOT, "minimap2-fpga_b200", "csrc", "chain_kernels.cu")


def kernel_sha():
    return hashlib.sha1(open(KERNEL_SRC, "rb").read()).hexdigest()[:12]


def committed_capture(workload, reads):
    """ncu --set full numbers of the dominant kernel from profiles/k1_capture.json — only if that capture was taken from THIS kernel
    source (sha of chain_kernels.cu) on this workload and size; otherwise None (a stale constant must not pass as a measurement)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "k1_capture.json")))
        if t["workload"] == workload and t["reads_per_gpu"] == reads and t["kernel_sha"] == kernel_sha():
            return t
    except Exception:
        pass
    return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def bind_near_gpu(torch, device_index):
    """Pin this rank's host threads (and therefore its first-touch pinned buffers) to the CPUs of the GPU's NUMA node, so that
    eight ranks do not push all their H2D/D2H traffic through one socket.  Best effort: silently skipped where sysfs says nothing."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (dom, bus, dev)
        cpus = open(path + "local_cpulist").read().strip()
        node = open(path + "numa_node").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        if ids and len(ids) < (os.cpu_count() or 1):
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": cpus}
    except Exception as e:      # noqa: BLE001
        return {"numa_node": None, "error": repr(e)[:80]}


def make_workload(name, n_reads, seed, source):
    """-> dict(off, a, par (oracle/binding keyword arguments), ref (recorded reference outputs or None), how, gen_s)."""
    t0 = time.time()
    if source == "real" and BW.available():
        w = BW.real_seed_batch(name, n_reads, seed)
        keys = ("max_dist_x", "max_dist_y", "bw", "max_skip", "max_iter", "min_cnt", "min_sc", "is_cdna", "n_segs", "gap_scale")
        par = dict(zip(keys, w["par"]))
        for k in keys[:-1]:
            par[k] = int(par[k])
        how = "anchors as the reference's own seeding hands them to mm_chain_dp (reference CLI `%s` on numpy-simulated reads, captured once, untimed: bench_workloads.py)" % w["meta"]["cli"]
        return dict(off=w["off"], a=w["a"], par=par, ref=w, how=how, gen_s=time.time() - t0)
    wl = load_package("workload")
    off, a = wl.preset_batch(name if name != "tandem" else "ultralong", n_reads, seed=seed)
    how = "anchors drawn from the seed-hit model in workload.py (calibrated against real minimap2 seeding, tests/golden/workload_calibration.json)"
    return dict(off=off, a=a, par={}, ref=None, how=how, gen_s=time.time() - t0)


def cpu_arm(w, n_threads, steps, warmup, budget_s=None):
    """Reference's CPU chaining over (a bounded sample of) the batch with n_threads host threads. Returns dict."""
    from oracle import oracle_py as O
    O.build()
    off, a = w["off"], w["a"]
    par = O.Params(**w["par"])
    n_reads = len(off) - 1
    ns = n_reads
    kind = "reference" if O.have_ref() else "port"
    if budget_s is not None and n_reads > 64:          # size the sample from a quick probe so that the arm stays within its budget
        probe = min(n_reads, max(64, n_reads // 50))
        t = O.replay(par, off[:probe + 1], a[:int(off[probe])], n_threads=n_threads, use_ref=(kind == "reference"), want_out=False)["seconds"]
        per_read = max(t / probe, 1e-9)
        ns = int(min(n_reads, max(probe, budget_s / (per_read * (steps + warmup)))))
    off_s, a_s = off[:ns + 1], a[:int(off[ns])]
    cells = O.replay(par, off_s, a_s, n_threads=n_threads, want_out=False)["stats"].cells      # the port counts cells; also warms caches
    times = []
    for it in range(warmup + steps):
        r = O.replay(par, off_s, a_s, n_threads=n_threads, use_ref=(kind == "reference"), want_out=False)
        if it >= warmup:
            times.append(r["seconds"])
    sec = sum(times) / len(times)
    return dict(value=cells / sec / 1e9, unit="GCUPS", cores=n_threads, kind=kind, seconds_per_step=sec, reads_per_s=ns / sec,
                sample="%d of %d reads (%d anchors, %d reference cells) per step, %d steps; inputs pre-copied outside the clock (the reference consumes a[])"
                       % (ns, n_reads, len(a_s), cells, steps),
                cells=int(cells), reads=ns)


def check_against_reference(res, w):
    """GPU results vs what the reference itself returned for the same calls when the workload was recorded: n_u, n_v, u[] and a hash of b[]."""
    ref = w["ref"]
    if ref is None:
        return None
    n_u, n_v = res["n_u"].astype(np.int64), res["n_v"].astype(np.int64)
    bad = int(np.count_nonzero(res["n_u"] != ref["ref_n_u"]) + np.count_nonzero(res["n_v"] != ref["ref_n_v"]))
    if bad == 0:
        idx = np.repeat(res["u_off"][:-1] - (np.cumsum(n_u) - n_u), n_u) + np.arange(int(n_u.sum()))
        bad += int(np.count_nonzero(res["u"][idx] != ref["ref_u"]))
        bw = res["b"].view(np.uint64) if "b" in res else None
        if bw is not None:
            for r in np.nonzero(n_v)[0][:20000]:        # hashes of b[] for the first 20k mapped reads (python loop: bounded)
                o = int(res["b_off"][r])
                bad += int(BW.b_hash(bw[2 * o:2 * (o + int(n_v[r]))]) != int(ref["ref_b_hash"][r]))
    return {"reads": int(len(n_u)), "mismatching_reads_or_entries": bad, "checked": "n_u, n_v, u[] of every read; hash of b[] for up to 20000 mapped reads; "
            "against the reference CLI's own outputs recorded with the workload"}


def device_value(torch, binding, w, local_rank, steps, warmup, barrier=None):
    """Throughput with the batch resident in HBM. -> dict(ms, k1_ms, cells, stats, launches)"""
    L = binding.load()
    par = binding.Params(**w["par"])
    db = binding.DeviceBatch(par, w["off"], w["a"], device=local_rank, index_out=True)
    db.set_counting(True)           # one untimed statistics pass: reference-semantics cells of this workload (the GCUPS numerator)
    db.run()
    st = db.stats()
    db.set_counting(False)
    for _ in range(warmup):
        db.run()
    launches0 = L.mm2b_launch_count()
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        db.run()
    e1.record()
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = L.mm2b_launch_count() - launches0
    k1 = []
    for _ in range(min(steps, 5)):   # dominant-kernel duration: CUDA events on the launching stream around that launch
        db.run()
        k1.append(db.chain_kernel_ms())
    res = db.results()
    st_timed = db.stats()
    db.close()
    return dict(ms=ms, k1_ms=sum(k1) / len(k1), cells=int(st.cells_ref), stats=st, stats_timed=st_timed, launches=launches, res=res)


def e2e_variants(torch, binding, w, steps, warmup, sync, which, keep_res=True):
    """Host-buffer batch calls on pinned arrays; -> {variant: dict(ms, stats, res)}"""
    off, a = w["off"], w["a"]
    n_reads, n_anchors = len(off) - 1, int(off[-1])
    par = binding.Params(**w["par"])
    pins = []

    def pin(n, dt):
        p = binding.PinnedArray(max(n, 1), dt)
        pins.append(p)
        return p.array

    out = {}
    try:
        h_a = pin(n_anchors, binding.ANCHOR)
        h_a[:n_anchors] = a
        bufs = {"u": pin(n_anchors, np.uint64), "n_u": pin(n_reads, np.int32), "n_v": pin(n_reads, np.int32), "status": pin(n_reads, np.int32),
                "u_off": pin(n_reads + 1, np.int64), "b_off": pin(n_reads + 1, np.int64)}
        h_b, h_bi = pin(n_anchors, binding.ANCHOR), pin(n_anchors, np.int32)
        modes = {"default": ("default", 0, True, False), "index": ("index", 0, False, True),
                 "host_gather": ("b", binding.F_HOST_GATHER, True, False),
                 "packed_b": ("b", binding.F_DEVICE_GATHER, True, False),
                 "raw_index": ("index", binding.F_RAW_INPUT, False, True)}
        for name in which:
            mode, flags, want_b, want_bi = modes[name]
            o = dict(bufs)
            if want_b:
                o["b"] = h_b
            if want_bi:
                o["bi"] = h_bi
            res = None
            for _ in range(warmup):
                res = binding.chain_batch(par, off, h_a[:n_anchors], out=o, mode=mode, flags=flags)
            sync()
            t0 = time.perf_counter()
            for _ in range(steps):
                res = binding.chain_batch(par, off, h_a[:n_anchors], out=o, mode=mode, flags=flags)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / steps
            out[name] = dict(ms=ms, stats=res["stats"], res={k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in res.items()} if name == "default" and keep_res else None)
    finally:
        for p in pins:
            p.free()
    return out


def _read_fingerprints(n_a, n_u, n_v, u_flat):
    n_u64 = np.asarray(n_u).astype(np.int64)
    h = np.zeros(len(n_u64), np.uint64)
    if len(u_flat):
        starts = np.cumsum(n_u64) - n_u64
        k = np.arange(len(u_flat), dtype=np.uint64) - np.repeat(starts, n_u64).astype(np.uint64)
        v = np.asarray(u_flat, np.uint64) * np.uint64(0x9E3779B97F4A7C15) + (k + np.uint64(1)) * np.uint64(0xD6E8FEB86659FD93)
        nz = np.flatnonzero(n_u64)
        h[nz] = np.add.reduceat(v, starts[nz])
    return h ^ (np.asarray(n_a).astype(np.uint64) << np.uint64(40)) ^ (n_u64.astype(np.uint64) << np.uint64(20)) ^ np.asarray(n_v).astype(np.uint64)


def frontend_e2e(torch, binding, name, n_reads, seed, w, steps, warmup, cells):
    """SURVEY 8(f) next-4 / next-1 in front of the chaining path: mm2b_map_batch — pinned ASCII read sequences in; sketch, seed hits, sort and
    chaining on the GPU; chains (u[], b[]), rep_len and mini_pos out.  Same reads as the anchor workload, so its chains are checked against
    the reference CLI's recorded outputs and its GCUPS counts the same reference cells."""
    fi = BW.front_inputs(name, n_reads, seed)
    if fi is None:
        return None
    t0 = time.perf_counter()
    idx = binding.Index(fi["index"])
    t_index = time.perf_counter() - t0
    pin = binding.PinnedArray(max(len(fi["seq"]), 1), np.uint8)
    try:
        pin.array[:len(fi["seq"])] = fi["seq"]
        par = binding.Params(**w["par"])
        res = None
        for _ in range(warmup):
            res = binding.map_batch(idx, None, fi["mid_occ"], par, seq_off=fi["seq_off"], blob=pin.array, collect="u")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = binding.map_batch(idx, None, fi["mid_occ"], par, seq_off=fi["seq_off"], blob=pin.array, collect=False)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        chk = binding.map_batch(idx, None, fi["mid_occ"], par, seq_off=fi["seq_off"], blob=pin.array, collect="u")
        ref = w["ref"]
        # the workload recorded the reference's calls in the order its threads made them, not in read order: compare the multisets of
        # per-read fingerprints (anchors seeded, n_u, n_v, a position-dependent hash of u[])
        fp_gpu = _read_fingerprints(chk["n_a"], chk["n_u"], chk["n_v"], chk["u_flat"])
        fp_ref = _read_fingerprints(np.diff(w["off"]), ref["ref_n_u"], ref["ref_n_v"], ref["ref_u"])
        bad = int(len(fp_gpu) != len(fp_ref)) or int(np.count_nonzero(np.sort(fp_gpu) != np.sort(fp_ref)))
        st = res["stats"]
        return {"api": "mm2b_map_batch (include/mm2seed_b200.h): pinned ASCII read sequences in; mm_sketch, seed hits, anchor sort and chaining on the GPU; u[], b[] (16 B), rep_len and "
                       "mini_pos out; host<->device copies inside the timed region",
                "value": cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms, "reads_per_s": n_reads / (ms * 1e-3), "bases_per_s": float(len(fi["seq"])) / (ms * 1e-3),
                "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]),
                "stage_ms_sum_over_subbatches": {"sketch": st["sketch_ms"], "seed": st["seed_ms"], "sort": st["sort_ms"], "chain": st["chain_ms"]},
                "minimizers": int(st["tot_mini"]), "anchors": int(st["tot_anchors"]), "reads_with_equal_keys": int(st["n_tie_reads"]), "segments": int(st["n_segs"]),
                "index": {"minimizers": int(len(fi["index"]["keys"])), "positions": int(len(fi["index"]["pos"])), "upload_and_build_s": t_index, "mid_occ": int(fi["mid_occ"])},
                "parity_check": {"reads": int(n_reads), "mismatching_reads_or_entries": bad,
                                 "checked": "multiset over reads of (anchors seeded, n_u, n_v, hash of u[]) against the reference CLI's own seeding + chaining recorded with the workload"}}
    finally:
        pin.free()
        idx.close()


def copy_ceiling(torch, devices, mbytes=512):
    """Pinned H2D and D2H bandwidth of this box, per GPU and with all `devices` copying at once (GB/s)."""
    n = mbytes << 20
    host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in devices]
    dev = [torch.empty(n, dtype=torch.uint8, device="cuda:%d" % d) for d in devices]
    streams = [torch.cuda.Stream(device="cuda:%d" % d) for d in devices]

    def run(direction, idx):
        for d in devices:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for _ in range(3):
            for i in idx:
                with torch.cuda.stream(streams[i]):
                    if direction == "h2d":
                        dev[i].copy_(host[i], non_blocking=True)
                    else:
                        host[i].copy_(dev[i], non_blocking=True)
        for d in devices:
            torch.cuda.synchronize(d)
        return 3 * len(idx) * n / (time.perf_counter() - t0) / 1e9

    run("h2d", [0])
    res = {"h2d_gbs_one_gpu": run("h2d", [0]), "d2h_gbs_one_gpu": run("d2h", [0])}
    if len(devices) > 1:
        allidx = list(range(len(devices)))
        res["h2d_gbs_all_gpus"], res["d2h_gbs_all_gpus"] = run("h2d", allidx), run("d2h", allidx)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="map-ont", choices=sorted(BW.PRESETS), help="map-ont is the configuration the metric is quoted on")
    ap.add_argument("--source", default="real", choices=["real", "model"], help="real = anchors from the reference CLI's seeding (default); model = round-1 anchor model")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (weak scaling); default 100000 / 50000 / 2000 / 96 by workload")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the asm20 / ultralong / tandem lines (N=1 only)")
    ap.add_argument("--no-frontend", action="store_true", help="skip the seeding front end line (N=1 only)")
    ap.add_argument("--strong-reads", type=int, default=0, help="additionally chain this many reads in ONE in-process call over all --gpus devices (BASELINE configs[4]: 1000000)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.reads <= 0:
        args.reads = BW.PRESETS[args.workload][3]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    def config_for(w, workload, reads):
        p = w["par"] or dict(max_dist_x=5000, max_dist_y=5000, bw=500, max_skip=25, max_iter=5000, min_cnt=3, min_sc=40)
        return {"workload": BW.DESCRIPTION[workload] % reads + "; " + w["how"] + "; chaining parameters of the preset (max_dist %d, bw %d, max_skip %d, max_iter %d, min_cnt %d, min_sc %d)"
                            % (p["max_dist_x"], p["bw"], p["max_skip"], p["max_iter"], p["min_cnt"], p["min_sc"]),
                "reads_per_gpu": reads, "l2": "inputs (anchors + 32 B/anchor scratch) exceed the 126 MB L2, no flush needed",
                "parallelism": "read-sharded, %d GPU(s), no collective" % world}

    # ---------------------------------------------------------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        w = make_workload(args.workload, args.reads, 1000, args.source)
        cores = os.cpu_count() or 1
        # each step = a bounded sample of the workload, sized so that the whole run stays within a few minutes
        r = cpu_arm(w, cores, max(args.steps, 1), max(args.warmup, 0), budget_s=150.0)
        line = {"impl": "reference", "metric": "chain_dp_gcups", "value": r["value"], "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic", "config": config_for(w, args.workload, args.reads), "reads_per_s": r["reads_per_s"],
                "cpu_baseline": {"value": r["value"], "unit": "GCUPS", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    binding = load_package("binding")
    L = binding.load()
    if not torch.cuda.is_available() or L.mm2b_cuda_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity = bind_near_gpu(torch, local_rank) if world > 1 else None
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")       # host-side waits that must not keep a kernel spinning on anybody's GPU
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w = make_workload(args.workload, args.reads, 1000 + rank, args.source)
    off, a = w["off"], w["a"]
    n_reads, n_anchors = len(off) - 1, int(off[-1])
    binding.init([local_rank])

    # ---- value: inputs resident in HBM -----------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dv = device_value(torch, binding, w, local_rank, args.steps, args.warmup, barrier=barrier)
    ms_value = max_over_ranks(dv["ms"])
    cells, st = dv["cells"], dv["stats"]
    tot_cells = sum_over_ranks(cells)
    tot_reads = sum_over_ranks(n_reads)
    tot_anchors = sum_over_ranks(n_anchors)
    value = tot_cells / (ms_value * 1e-3) / 1e9

    # ---- e2e: pinned host buffers through the host-buffer batch call, one process per GPU ------------------------
    ev = e2e_variants(torch, binding, w, args.steps, args.warmup, barrier, ["default", "index", "raw_index", "host_gather", "packed_b"] if world == 1 else ["default", "index", "raw_index"])
    ms_e2e = {k: max_over_ranks(v["ms"]) for k, v in ev.items()}
    clocks = sampler.stop() if rank == 0 else None      # sampled across both timed regions (value and e2e)
    est = ev["default"]["stats"]
    h2d, d2h = sum_over_ranks(int(est.h2d_bytes)), sum_over_ranks(int(est.d2h_bytes))

    def e2e_obj(name, api):
        s = ev[name]["stats"]
        return {"value": tot_cells / (ms_e2e[name] * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms_e2e[name], "reads_per_s": tot_reads / (ms_e2e[name] * 1e-3), "api": api,
                "h2d_bytes_per_step": int(sum_over_ranks(int(s.h2d_bytes))), "d2h_bytes_per_step": int(sum_over_ranks(int(s.d2h_bytes))),
                "stage_ms_sum_over_subbatches_rank0": {"h2d": s.h2d_ms, "kernels": s.kernel_ms, "d2h": s.d2h_ms, "host_pack": s.pack_ms, "host_gather": s.gather_ms},
                "subbatches_rank0": {"packed": int(s.n_packed_subs), "raw": int(s.n_raw_subs)}}

    apis = {"default": "mm2b_chain_batch: pinned mm128_t anchors in (16 B/anchor over PCIe), u[] and b[] (mm128_t, gathered on the device) out",
            "index": "mm2b_chain_batch_ex with bi[] (what the product's own callers use: mm_chain_dp's batcher and the phase-split caller hold a[]): input packed to 8 B/anchor by the "
                     "library's helper threads, chained anchors back as int32 indices",
            "host_gather": "mm2b_chain_batch_ex(MM2B_F_HOST_GATHER): packed input, int32 indices over PCIe, b[] gathered from the caller's a[] by the library's helper threads",
            "packed_b": "mm2b_chain_batch_ex(MM2B_F_DEVICE_GATHER): packed input, b[] as 16-byte anchors from the device",
            "raw_index": "mm2b_chain_batch_ex(MM2B_F_RAW_INPUT) with bi[]: 16 B/anchor in as it is (no host pass at all), int32 indices out: the least host memory traffic per anchor"}
    e2e = e2e_obj("default", apis["default"])
    e2e["variants"] = {k: e2e_obj(k, apis[k]) for k in ev if k != "default"}
    e2e["host_threads"] = "helper pool of the library (MM2B_HOST_THREADS, default min(cores-2, 16)) + 1 worker per device"
    parity = check_against_reference(ev["default"]["res"], w) if rank == 0 else None
    ev = None

    # ---- the library's own multi-GPU path: ONE call from rank 0 over all N devices (N > 1) --------------------------
    multi = None
    if world > 1:
        binding.shutdown()
        torch.cuda.synchronize(dev)
        dist.barrier(group=host_group)
        if rank == 0:
            ws = [w] + [make_workload(args.workload, args.reads, 1000 + r, args.source) for r in range(1, world)]     # the other ranks' batches (cached on disk by them)
            big_off = np.concatenate([[0]] + [x["off"][1:] + sum(int(y["off"][-1]) for y in ws[:i]) for i, x in enumerate(ws)]).astype(np.int64)
            big = dict(off=big_off, a=np.concatenate([x["a"] for x in ws]), par=w["par"])
            ws = None
            binding.init(list(range(world)))
            ceiling = copy_ceiling(torch, list(range(world)))
            mv = e2e_variants(torch, binding, big, args.steps, args.warmup, lambda: None, ["default", "index", "raw_index"], keep_res=False)
            binding.shutdown()
            cells_all = tot_cells
            multi = {k: {"value": cells_all / (v["ms"] * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": v["ms"], "reads_per_s": (len(big_off) - 1) / (v["ms"] * 1e-3),
                         "h2d_bytes_per_step": int(v["stats"].h2d_bytes), "d2h_bytes_per_step": int(v["stats"].d2h_bytes),
                         "stage_ms_sum_over_subbatches": {"h2d": v["stats"].h2d_ms, "kernels": v["stats"].kernel_ms, "d2h": v["stats"].d2h_ms,
                                                          "host_pack": v["stats"].pack_ms, "host_gather": v["stats"].gather_ms}} for k, v in mv.items()}
            multi["copy_ceiling"] = ceiling
            big = None
        dist.barrier(group=host_group)
        binding.init([local_rank])

    # ---- CPU baseline on rank 0, N=1 only -------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(w, os.cpu_count() or 1, steps=2, warmup=1, budget_s=20.0)
        cpu = {"value": r["value"], "unit": "GCUPS", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "reads_per_s": r["reads_per_s"]}
        if r["reads"] == n_reads:
            assert r["cells"] == cells, "GPU cell tally %d != oracle %d" % (cells, r["cells"])

    # ---- the seeding front end in front of the same reads (N=1 only) ------------------------------------------------------
    front = None
    if rank == 0 and world == 1 and args.workload == "map-ont" and w["ref"] is not None and not args.no_frontend:
        try:
            front = frontend_e2e(torch, binding, args.workload, args.reads, 1000, w, min(args.steps, 5), 2, cells)
        except Exception as e:      # noqa: BLE001
            front = {"error": repr(e)[:300]}

    # ---- the other configurations of BASELINE.json, inside the same line (N=1 only) -------------------------------
    others = None
    if rank == 0 and world == 1 and not args.no_other_workloads and args.workload == "map-ont":
        others = {}
        for name in ("asm20", "ultralong", "tandem"):
            try:
                reads_o = BW.PRESETS[name][3]
                wo = make_workload(name, reads_o, 1000, args.source)
                d = device_value(torch, binding, wo, local_rank, steps=min(args.steps, 5), warmup=3)
                eo = e2e_variants(torch, binding, wo, min(args.steps, 5), 3, lambda: torch.cuda.synchronize(), ["default"])["default"]
                par_o = check_against_reference(eo["res"], wo)
                co = cpu_arm(wo, os.cpu_count() or 1, steps=1, warmup=0, budget_s=6.0) if not args.no_cpu_baseline else None
                n = np.diff(wo["off"])
                others[name] = {"workload": BW.DESCRIPTION[name] % reads_o, "reads": int(len(n)), "anchors": int(wo["off"][-1]), "anchors_per_read_mean": float(n.mean()),
                                "anchors_per_read_max": int(n.max()), "cells_per_step": d["cells"], "value": d["cells"] / (d["ms"] * 1e-3) / 1e9, "unit": "GCUPS",
                                "ms_per_step": d["ms"], "kernel_ms": d["k1_ms"], "reads_per_s": len(n) / (d["ms"] * 1e-3), "n_heavy_reads": int(eo["stats"].n_heavy_reads),
                                "e2e": {"value": d["cells"] / (eo["ms"] * 1e-3) / 1e9, "ms_per_step": eo["ms"], "h2d_bytes_per_step": int(eo["stats"].h2d_bytes),
                                        "d2h_bytes_per_step": int(eo["stats"].d2h_bytes)},
                                "parity_check": par_o, "cpu_baseline": None if co is None else {"value": co["value"], "cores": co["cores"], "kind": co["kind"], "sample": co["sample"]}}
                if name == "ultralong":
                    # the same recorded reads five times over (10,000 reads, 71 M anchors): what the kernel does on ultra-long reads once the
                    # batch fills the GPU — 2,000 reads occupy 2,000 of its 4,736 warp slots and the launch lasts as long as its longest read
                    rep = 5
                    n_a = int(wo["off"][-1])
                    off5 = np.concatenate([[0]] + [wo["off"][1:] + k * n_a for k in range(rep)]).astype(np.int64)
                    w5 = dict(off=off5, a=np.tile(wo["a"], rep), par=wo["par"], ref=None)
                    d5 = device_value(torch, binding, w5, local_rank, steps=min(args.steps, 5), warmup=3)
                    others["ultralong_x5"] = {"workload": "the ultra-long batch above five times over: %d reads, %d anchors per step" % (len(off5) - 1, int(off5[-1])),
                                              "cells_per_step": d5["cells"], "value": d5["cells"] / (d5["ms"] * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": d5["ms"],
                                              "kernel_ms": d5["k1_ms"], "reads_per_s": (len(off5) - 1) / (d5["ms"] * 1e-3)}
                    w5 = d5 = None
                    # long-read segmenting: eight recorded reads at a time merged into one chimeric read (their anchors together, sorted by x,
                    # the query coordinates of each shifted behind the previous one's): the loci are x-gap cut points (chain.c:192), the library
                    # cuts such reads into pieces filled by warps of their own.  Checked against the oracle; timed with and without cutting.
                    grp = 8
                    parts, offs = [], [0]
                    for g0 in range(0, len(n) - grp + 1, grp):
                        qshift, chunk = 0, []
                        for r in range(g0, g0 + grp):
                            ar = wo["a"][wo["off"][r]:wo["off"][r + 1]].copy()
                            ar["y"] += np.uint64(qshift)
                            qshift += 1 << 20
                            chunk.append(ar)
                        ch = np.concatenate(chunk)
                        parts.append(ch[np.argsort(ch["x"], kind="stable")])
                        offs.append(offs[-1] + len(ch))
                    wc = dict(off=np.asarray(offs, np.int64), a=np.concatenate(parts), par=wo["par"], ref=None)
                    dc = device_value(torch, binding, wc, local_rank, steps=min(args.steps, 5), warmup=3)
                    os.environ["MM2B_SEG"] = "0"
                    try:
                        dn = device_value(torch, binding, wc, local_rank, steps=min(args.steps, 5), warmup=3)
                    finally:
                        os.environ.pop("MM2B_SEG", None)
                    from oracle import oracle_py as O
                    oref = O.replay(O.Params(**wc["par"]), wc["off"], wc["a"], n_threads=os.cpu_count() or 1)
                    rc_ = dc["res"]
                    bad = int(np.count_nonzero(rc_["n_u"] != oref["n_u"]) + np.count_nonzero(rc_["n_v"].astype(np.int64) != oref["n_v"].astype(np.int64)))
                    if bad == 0:
                        for r in range(len(offs) - 1):
                            o_, nu_ = int(offs[r]), int(oref["n_u"][r])
                            bad += int(not np.array_equal(rc_["u"][int(rc_["u_off"][r]):int(rc_["u_off"][r]) + nu_], oref["u"][o_:o_ + nu_]))
                    others["ultralong_chimeric"] = {"workload": "the ultra-long reads merged eight at a time into chimeric reads: %d reads, %d anchors (mean %d per read)"
                                                                % (len(offs) - 1, int(offs[-1]), int(offs[-1]) // max(1, len(offs) - 1)),
                                                    "cells_per_step": dc["cells"], "value": dc["cells"] / (dc["ms"] * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": dc["ms"],
                                                    "reads_cut_into_pieces": int(dc["stats_timed"].n_cut_reads), "ms_per_step_without_cutting": dn["ms"],
                                                    "value_without_cutting": dn["cells"] / (dn["ms"] * 1e-3) / 1e9,
                                                    "parity_check": {"reads": len(offs) - 1, "mismatching_reads_or_entries": bad, "checked": "n_u, n_v, u[] of every read against the oracle"}}
                    wc = dc = dn = oref = None
            except Exception as e:      # noqa: BLE001
                others[name] = {"error": repr(e)[:300]}

    # ---- BASELINE configs[4]: strong scaling of one big in-process call (opt-in) ------------------------------------------
    strong = None
    if args.strong_reads > 0 and rank == 0:
        n_batches = (args.strong_reads + args.reads - 1) // args.reads
        ws = [make_workload(args.workload, args.reads, 1000 + r, args.source) for r in range(n_batches)]
        big_off = np.concatenate([[0]] + [x["off"][1:] + sum(int(y["off"][-1]) for y in ws[:i]) for i, x in enumerate(ws)]).astype(np.int64)
        big = dict(off=big_off, a=np.concatenate([x["a"] for x in ws]), par=w["par"])
        ws = None
        binding.shutdown()
        strong = {"reads": int(len(big_off) - 1), "anchors": int(big_off[-1]), "devices": {}}
        for nd in [n for n in (1, 2, 4, 8) if n <= args.gpus]:
            binding.init(list(range(nd)))
            binding.set_counting(True)
            cells_big = int(binding.chain_batch(binding.Params(**big["par"]), big["off"], big["a"], mode="index")["stats"].cells_ref)
            binding.set_counting(False)
            v = e2e_variants(torch, binding, big, max(2, min(args.steps, 5)), 2, lambda: None, ["default"], keep_res=False)["default"]
            strong["devices"][str(nd)] = {"gcups": cells_big / (v["ms"] * 1e-3) / 1e9, "ms": v["ms"], "reads_per_s": (len(big_off) - 1) / (v["ms"] * 1e-3)}
            binding.shutdown()
        binding.init([local_rank])

    if rank == 0:
        peaks, how = measured_peaks()
        int_peak_instr = L.mm2b_measure_int32_peak(local_rank)      # G lane-instructions/s, measured live on this GPU
        k1_avg = dv["k1_ms"]
        hbm_ach = BYTES_PER_ANCHOR * n_anchors / (k1_avg * 1e-3) / 1e9
        int_ach = cells * INT_OPS_PER_CELL / (k1_avg * 1e-3) / 1e9
        cap = committed_capture(args.workload, args.reads)
        line = {"metric": "chain_dp_gcups", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": config_for(w, args.workload, args.reads), "reads_per_s": tot_reads / (ms_value * 1e-3), "anchors_per_s": tot_anchors / (ms_value * 1e-3),
                "cells_per_step": int(tot_cells), "cells_issued_per_step_rank0": int(st.cells_issued), "window_cells_per_step_rank0": int(st.window_cells), "anchors_per_step": int(tot_anchors),
                "e2e": e2e, "gpu_launches": int(dv["launches"]), "clocks": clocks, "parity_check": parity,
                "roofline": {"bound": "hbm", "kernel": "chain_reads_kernel", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": hbm_ach / peaks["hbm_gbs"], "traffic": (cap["dram_bytes_read_per_launch"] + cap["dram_bytes_write_per_launch"]) if cap else None,
                             "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)", "algorithmic_bytes_per_launch": BYTES_PER_ANCHOR * n_anchors, "kernel_ms": k1_avg,
                             "note": "the path is INT32-issue bound, not HBM bound (SURVEY.md 8d): see roofline_int32"},
                "roofline_int32": {"bound": "int32_issue", "achieved": int_ach, "unit": "Gop/s", "ops_per_cell": INT_OPS_PER_CELL,
                                   "peak_lane_instr_per_s": int_peak_instr, "peak": int_peak_instr * 1.5, "frac": int_ach / (int_peak_instr * 1.5) if int_peak_instr > 0 else None,
                                   "frac_uses": "peak = 1.5 x the measured lane-instruction peak: the micro-benchmark issues LOP3 + VIADDMNMX (2 SASS instructions) per statement of 3 "
                                                "integer operations (profiles/int32_peak_sass.txt); achieved = 30 algorithmic ops x reference cells / kernel time",
                                   "ncu_capture": None if not cap else {k: cap[k] for k in ("issue_slots_busy_pct", "alu_pipe_pct", "fma_pipe_pct", "warp_instr_per_anchor", "file") if k in cap},
                                   "peak_source": "measured live: mm2b_measure_int32_peak"},
                "cpu_baseline": cpu, "frontend": front, "other_workloads": others, "e2e_per_process": None, "strong_scaling": strong, "workload_gen_s": w["gen_s"], "host_affinity_rank0": affinity}
        if multi is not None:        # N > 1: the headline e2e is the library's own multi-device call; the one-process-per-GPU number stays beside it
            line["e2e_per_process"] = e2e
            m = multi["default"]
            line["e2e"] = {"value": m["value"], "unit": "GCUPS", "ms_per_step": m["ms_per_step"], "reads_per_s": m["reads_per_s"], "h2d_bytes_per_step": m["h2d_bytes_per_step"],
                           "d2h_bytes_per_step": m["d2h_bytes_per_step"], "api": "ONE mm2b_chain_batch call from rank 0 over all %d devices (per-device worker threads inside the library) on %d reads"
                           % (world, int(tot_reads)), "stage_ms_sum_over_subbatches": m["stage_ms_sum_over_subbatches"], "variants": {k: multi[k] for k in ("index", "raw_index") if k in multi},
                           "copy_ceiling_GBps": multi["copy_ceiling"]}
        print(json.dumps(line), flush=True)
    binding.shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
