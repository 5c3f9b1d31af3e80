"""Do copies in opposite directions overlap on this box when every pipeline stream carries both?  python tools/copy_probe.py
(a) 6 streams, each alternating H2D 64 MB / D2H 96 MB; (b) the same bytes, every H2D on one stream and every D2H on another;
(c) 6 streams, three of them H2D only and three D2H only.  Prints the time of each and what one direction alone takes."""
import time, torch
MB = 1 << 20
n_in, n_out, rounds, ns = 64 * MB, 96 * MB, 5, 6
hin = [torch.empty(n_in, dtype=torch.uint8).pin_memory() for _ in range(ns)]
hout = [torch.empty(n_out, dtype=torch.uint8).pin_memory() for _ in range(ns)]
din = [torch.empty(n_in, dtype=torch.uint8, device="cuda") for _ in range(ns)]
dout = [torch.empty(n_out, dtype=torch.uint8, device="cuda") for _ in range(ns)]
streams = [torch.cuda.Stream() for _ in range(ns)]
def run(plan):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plan()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3
def both_per_stream():
    for r in range(rounds):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                din[i].copy_(hin[i], non_blocking=True)
                hout[i].copy_(dout[i], non_blocking=True)
def two_streams():
    for r in range(rounds):
        for i in range(ns):
            with torch.cuda.stream(streams[0]): din[i].copy_(hin[i], non_blocking=True)
            with torch.cuda.stream(streams[1]): hout[i].copy_(dout[i], non_blocking=True)
def split_streams():
    for r in range(rounds):
        for i in range(ns):
            with torch.cuda.stream(streams[i % 3]): din[i].copy_(hin[i], non_blocking=True)
            with torch.cuda.stream(streams[3 + i % 3]): hout[i].copy_(dout[i], non_blocking=True)
def only_in():
    for r in range(rounds):
        for i in range(ns):
            with torch.cuda.stream(streams[i]): din[i].copy_(hin[i], non_blocking=True)
def only_out():
    for r in range(rounds):
        for i in range(ns):
            with torch.cuda.stream(streams[i]): hout[i].copy_(dout[i], non_blocking=True)
gb_in, gb_out = rounds * ns * n_in / 1e9, rounds * ns * n_out / 1e9
for name, plan in (("H2D alone", only_in), ("D2H alone", only_out), ("(a) both directions on each of 6 streams", both_per_stream),
                   ("(b) one H2D stream + one D2H stream", two_streams), ("(c) 3 H2D streams + 3 D2H streams", split_streams)):
    run(plan)
    ms = min(run(plan) for _ in range(3))
    print("%-45s %7.2f ms  (H2D %.2f GB, D2H %.2f GB; if serial at 55 GB/s: %.1f ms)" % (name, ms, gb_in, gb_out, (gb_in + gb_out) / 55 * 1e3), flush=True)
