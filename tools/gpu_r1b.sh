#!/bin/bash
# round-1 second GPU session: parity suite incl. tcgen05 modes, bench in all three conv modes, launch lists
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for mode in bf16x3 bf16 fp32; do
  timeout 600 python bench.py --steps 5 --warmup 3 --conv-mode $mode > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  cat gpurun_out/bench_$mode.json; tail -n 2 gpurun_out/bench_$mode.err
done
timeout 600 python bench.py --workload pipeline_ref --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pipeline_bf16x3.json 2> gpurun_out/bench_pipeline.err
cat gpurun_out/bench_pipeline_bf16x3.json; tail -n 2 gpurun_out/bench_pipeline.err
for mode in bf16x3 bf16; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$mode.csv \
    python bench.py --steps 1 --warmup 3 --batch 4 --no-cpu-baseline --conv-mode $mode > gpurun_out/ncu_launch_$mode.log 2>&1
done
