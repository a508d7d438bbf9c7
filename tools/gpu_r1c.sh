#!/bin/bash
# plane-layout + TMA/tcgen05 conv kernel bring-up: parity suite, per-tap debug, bench
mkdir -p gpurun_out
timeout 300 python tools/tc_debug.py REF > gpurun_out/tc_debug_ref.log 2>&1; echo "rc=$?" >> gpurun_out/tc_debug_ref.log
tail -12 gpurun_out/tc_debug_ref.log
timeout 300 python tools/tc_debug.py PAPER > gpurun_out/tc_debug_paper.log 2>&1; echo "rc=$?" >> gpurun_out/tc_debug_paper.log
tail -8 gpurun_out/tc_debug_paper.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
for mode in bf16x3 bf16; do
  timeout 600 python bench.py --steps 5 --warmup 3 --conv-mode $mode --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  cat gpurun_out/bench_$mode.json; tail -n 2 gpurun_out/bench_$mode.err
done
