#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_trace.py bf16x3 160 127 > gpurun_out/trace_rs_160_127.log 2>&1
timeout 300 python tools/tc_trace.py bf16x3 120 255 > gpurun_out/trace_rs_120_255.log 2>&1
head -3 gpurun_out/trace_rs_160_127.log; head -3 gpurun_out/trace_rs_120_255.log
