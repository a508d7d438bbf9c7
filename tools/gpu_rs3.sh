#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_trace.py bf16x3 160 127 > gpurun_out/trace_rs_160_127.log 2>&1
timeout 300 python tools/tc_trace.py bf16 160 127 > gpurun_out/trace_rs_160_127_bf16.log 2>&1
for G in 4 5; do
MISO_RS_G=$G timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_G$G.log 2>&1; tail -1 gpurun_out/lt_rs_G$G.log
done
