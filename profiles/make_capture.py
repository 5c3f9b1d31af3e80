#!/usr/bin/env python
"""profiles/k1_capture.json from an `ncu --set full` report of chain_reads_kernel: the numbers bench.py quotes next to its live timings
(DRAM bytes per launch, issue-slot / pipe utilisation, warp-instructions per anchor), stamped with the sha of the kernel source they were
taken from — bench.py ignores the file when the source has changed since.
   python profiles/make_capture.py <report.ncu-rep> <workload> <reads_per_gpu> <anchors_per_launch> <summary file under profiles/>"""
import csv, hashlib, io, json, os, subprocess, sys
rep, workload, reads, anchors, summary = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
here = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name, scale=1.0):
    i = hdr.index(name)
    u = units[i]
    f = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
    return sum(float(r[i].replace(",", "")) for r in data) / len(data) * f * scale
src = os.path.join(here, "..", "minimap2-fpga_b200", "csrc", "chain_kernels.cu")
out = {"workload": workload, "reads_per_gpu": reads, "kernel_sha": hashlib.sha1(open(src, "rb").read()).hexdigest()[:12], "launches_captured": len(data),
       "kernel_ms_under_ncu": col("gpu__time_duration.sum"),
       "dram_bytes_read_per_launch": int(col("dram__bytes_read.sum")), "dram_bytes_write_per_launch": int(col("dram__bytes_write.sum")),
       "issue_slots_busy_pct": round(col("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
       "alu_pipe_pct": round(col("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 2),
       "fma_pipe_pct": round(col("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2),
       "warp_instr_per_launch": int(col("smsp__inst_executed.sum")), "warp_instr_per_anchor": round(col("smsp__inst_executed.sum") / anchors, 2),
       "file": "profiles/" + os.path.basename(summary)}
json.dump(out, open(os.path.join(here, "k1_capture.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
