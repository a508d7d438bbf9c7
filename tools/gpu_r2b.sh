#!/bin/bash
# round 2, call B: fused operand preparation of the DenseBlock convs -- GPU tests, bench, per-layer times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log; tail -25 gpurun_out/r2b_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err; cut -c1-400 gpurun_out/r2b_bench_default.json; tail -2 gpurun_out/r2b_bench_default.err
timeout 300 python tools/layer_times.py > gpurun_out/r2b_layer_times.log 2>&1; tail -3 gpurun_out/r2b_layer_times.log
MISO_RS_FORK=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_nofork.json 2> gpurun_out/r2b_bench_nofork.err; cut -c1-300 gpurun_out/r2b_bench_nofork.json
MISO_RS_FUSE=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_nofuse.json 2> gpurun_out/r2b_bench_nofuse.err; cut -c1-300 gpurun_out/r2b_bench_nofuse.json
timeout 120 ./tools/umma_mn_test > gpurun_out/r2b_umma_mn_shift.log 2>&1; tail -20 gpurun_out/r2b_umma_mn_shift.log
