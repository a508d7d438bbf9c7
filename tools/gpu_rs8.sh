#!/bin/bash
mkdir -p gpurun_out
MISO_TC_DEBUG=1 timeout 900 python -m pytest tests -m gpu -q -x -k "net" -s > gpurun_out/pytest_rs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rs.log; grep -v "^conv_\|^$" gpurun_out/pytest_rs.log | tail -15; grep "conv_rs.*S=[24]" gpurun_out/pytest_rs.log | sort | uniq | head -6
timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_bf16x3.log 2>&1; tail -1 gpurun_out/lt_rs_bf16x3.log
timeout 300 python tools/layer_times.py bf16 > gpurun_out/lt_rs_bf16.log 2>&1; tail -1 gpurun_out/lt_rs_bf16.log
