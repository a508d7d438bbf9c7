#!/bin/bash
# first run of the row-streaming conv kernel: parity tests of the net, layer times old/new, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "net" > gpurun_out/pytest_rs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rs.log; tail -15 gpurun_out/pytest_rs.log
MISO_TC_DEBUG=1 timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_bf16x3.log 2> gpurun_out/lt_rs_bf16x3.err; grep conv_rs gpurun_out/lt_rs_bf16x3.err | sort | uniq | head -20; tail -3 gpurun_out/lt_rs_bf16x3.log
timeout 300 python tools/layer_times.py bf16 > gpurun_out/lt_rs_bf16.log 2>&1; tail -1 gpurun_out/lt_rs_bf16.log
timeout 600 python bench.py > gpurun_out/bench_rs.json 2> gpurun_out/bench_rs.err; cut -c1-400 gpurun_out/bench_rs.json; tail -2 gpurun_out/bench_rs.err
