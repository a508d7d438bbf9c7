#!/bin/bash
# round 2, call K: where does the conv_rs epilogue spend its time?  cycle traces with per-item tags, layer times with parts disabled
mkdir -p gpurun_out
export MISO_TRACE_KERNEL=rs
timeout 200 python tools/tc_trace.py bf16x3 32 127 > gpurun_out/r2k_trace_cin32_f127.log 2>&1
timeout 200 python tools/tc_trace.py bf16x3 160 127 > gpurun_out/r2k_trace_cin160_f127.log 2>&1
for d in 0 1 2 4 8 16 31; do
  MISO_RS_DBG=$d timeout 200 python tools/layer_times.py > gpurun_out/r2k_layer_times_dbg$d.log 2>&1
  echo "dbg $d: $(tail -1 gpurun_out/r2k_layer_times_dbg$d.log) | $(awk '$1==3||$1==5||$1==15||$1==17||$1==23||$1==27||$1==39||$1==51{printf "%s:%s ", $1, $3}' gpurun_out/r2k_layer_times_dbg$d.log)"
done
