#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/umma_bench > gpurun_out/umma_bench.log 2>&1; cat gpurun_out/umma_bench.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
