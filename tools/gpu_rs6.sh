#!/bin/bash
mkdir -p gpurun_out
MISO_TC_DEBUG=1 MISO_RS_MAXNC=64 timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_nc64.log 2> gpurun_out/lt_rs_nc64.err; grep "conv_rs.*cout=\(48\|64\)" gpurun_out/lt_rs_nc64.err | sort | uniq; sed -n '92p;98p;100p' gpurun_out/lt_rs_nc64.log
MISO_RS_MAXNC=64 timeout 600 python -m pytest tests -m gpu -q -x -k "net" > gpurun_out/pytest_rs.log 2>&1; tail -3 gpurun_out/pytest_rs.log
