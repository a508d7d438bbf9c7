#!/bin/bash
# round 2, call O: the TCN as one cluster-per-sample launch -- parity (tcn tap), bench A/B
mkdir -p gpurun_out
T=${1:-r2o}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest_parity.log 2>&1; tail -6 gpurun_out/${T}_pytest_parity.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-330 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
MISO_TCN_FUSED=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_unfused.json 2> gpurun_out/${T}_bench_unfused.err; cut -c1-330 gpurun_out/${T}_bench_unfused.json
