#!/bin/bash
mkdir -p gpurun_out
MISO_TC_DEBUG=1 timeout 900 python -m pytest tests -m gpu -q -x -k "net" -s > gpurun_out/pytest_rs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rs.log; grep -v "^conv_\|^$" gpurun_out/pytest_rs.log | tail -12; grep "conv_rs.*S=\(8\|16\)" gpurun_out/pytest_rs.log | sort | uniq | head -4
timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_bf16x3.log 2>&1; tail -1 gpurun_out/lt_rs_bf16x3.log
MISO_RS_PACKED_MINF=15 timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_bf16x3_min15.log 2>&1; tail -1 gpurun_out/lt_rs_bf16x3_min15.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; cut -c1-260 gpurun_out/bench_x.json; tail -2 gpurun_out/bench_x.err
timeout 600 python bench.py --workload miso1_ref --no-cpu-baseline > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; cut -c1-260 gpurun_out/bench_y.json
