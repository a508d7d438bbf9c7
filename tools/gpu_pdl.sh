#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "net or pipeline" > gpurun_out/pytest_rs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rs.log; tail -6 gpurun_out/pytest_rs.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err; cut -c1-260 gpurun_out/bench_pdl1.json; tail -2 gpurun_out/bench_pdl1.err
MISO_PDL=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; cut -c1-260 gpurun_out/bench_pdl0.json
timeout 300 python tools/determinism.py > gpurun_out/determinism.log 2>&1; tail -3 gpurun_out/determinism.log
