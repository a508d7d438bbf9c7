#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_trace.py bf16x3 32 63 > gpurun_out/trace_rs_32_63.log 2>&1
timeout 300 python tools/tc_trace.py bf16x3 24 255 > gpurun_out/trace_rs_24_255.log 2>&1
head -3 gpurun_out/trace_rs_32_63.log | cut -c1-700
