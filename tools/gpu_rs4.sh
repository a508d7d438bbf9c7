#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "net" > gpurun_out/pytest_rs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rs.log; tail -4 gpurun_out/pytest_rs.log
timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_bf16x3.log 2>&1; tail -1 gpurun_out/lt_rs_bf16x3.log
timeout 300 python tools/layer_times.py bf16 > gpurun_out/lt_rs_bf16.log 2>&1; tail -1 gpurun_out/lt_rs_bf16.log
timeout 300 python tools/tc_trace.py bf16x3 160 127 > gpurun_out/trace_rs_160_127.log 2>&1
timeout 300 python tools/tc_trace.py bf16 160 127 > gpurun_out/trace_rs_160_127_bf16.log 2>&1
