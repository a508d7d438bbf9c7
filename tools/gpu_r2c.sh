#!/bin/bash
mkdir -p gpurun_out
MISO_TRACE_KERNEL=rs python tools/tc_trace.py bf16x3 24 255 > gpurun_out/r2c_trace_conv1.log 2>&1
MISO_TRACE_KERNEL=rs python tools/tc_trace.py bf16x3 72 255 > gpurun_out/r2c_trace_conv3.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; cut -c1-300 gpurun_out/r2c_bench_default.json; tail -2 gpurun_out/r2c_bench_default.err
timeout 300 python tools/layer_times.py > gpurun_out/r2c_layer_times.log 2>&1; tail -1 gpurun_out/r2c_layer_times.log
MISO_RS_FUSE=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_nofuse.json 2> gpurun_out/r2c_bench_nofuse.err; cut -c1-300 gpurun_out/r2c_bench_nofuse.json
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2c_pytest_gpu.log
