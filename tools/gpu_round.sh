#!/bin/bash
# One GPU-box visit: GPU tests, bench lines, launch list and a full ncu capture of the dominant kernel. Everything lands in gpurun_out/<tag>/.
# usage: tools/gpu_round.sh <tag> [tests|notests]
tag=${1:-r2a}; mode=${2:-tests}
out=gpurun_out/$tag; mkdir -p $out
export CUDA_DEVICE_MAX_CONNECTIONS=16
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
nproc >> $out/gpu.txt
if [ "$mode" = tests ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
  echo "pytest exit $?" >> $out/pytest.log
fi
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_reference.log 2>&1
( time timeout 1500 python bench.py ) > $out/bench.log 2>&1
rc=$?; echo "bench exit $rc" >> $out/bench.log
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-workloads > $out/ncu_launches.log 2>&1
  [ -n "$SKIP_FULL" ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_reads_kernel -s 4 -c 2 -o $out/chain_reads_full -f \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-workloads > $out/ncu_full.log 2>&1
fi
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
( MM2B_MAP_TRACE=1 timeout 120 python tools/front_quick.py 100000 6 ) > $out/front_quick_6ctx.txt 2>&1
( MM2B_MAP_RAMP=0 timeout 120 python tools/front_quick.py 100000 6 | tail -1 ) > $out/front_quick_6ctx_noramp.txt 2>&1
tail -1 $out/front_quick_6ctx.txt | cut -c1-150; tail -1 $out/front_quick_6ctx_noramp.txt | cut -c1-150; tail -2 $out/smoke.log
tail -3 $out/pytest.log 2>/dev/null; tail -2 $out/bench.log | cut -c1-600
