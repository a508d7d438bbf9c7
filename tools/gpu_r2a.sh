#!/bin/bash
# round 2, call A: CTA-pair MMA probe, GPU tests, default bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
timeout 120 ./tools/umma_2cta_test > gpurun_out/r2a_umma_2cta.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_umma_2cta.log; cat gpurun_out/r2a_umma_2cta.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log; tail -15 gpurun_out/r2a_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; cut -c1-400 gpurun_out/r2a_bench_default.json; tail -2 gpurun_out/r2a_bench_default.err
