#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "net_tensor_core or row_streaming" > gpurun_out/pytest_rs.log 2>&1; tail -2 gpurun_out/pytest_rs.log
bash tools/gpu_lt.sh
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; cut -c1-260 gpurun_out/bench_x.json; tail -2 gpurun_out/bench_x.err
