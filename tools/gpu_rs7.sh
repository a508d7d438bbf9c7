#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "net" > gpurun_out/pytest_rs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rs.log; tail -12 gpurun_out/pytest_rs.log
timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_bf16x3.log 2>&1; tail -1 gpurun_out/lt_rs_bf16x3.log
timeout 600 python bench.py > gpurun_out/bench_rs.json 2> gpurun_out/bench_rs.err; cut -c1-330 gpurun_out/bench_rs.json; tail -2 gpurun_out/bench_rs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rs_bf16x3.csv python tools/one_fwd.py bf16x3 2 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
