#!/bin/bash
# round 2, call L: new conv_rs epilogue (chunk-per-warp split, pipelined accumulator loads, packed fp32 math) -- parity, layer times, bench
mkdir -p gpurun_out
T=${1:-r2l}
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest_parity.log 2>&1; tail -4 gpurun_out/${T}_pytest_parity.log
timeout 300 python tools/layer_times.py > gpurun_out/${T}_layer_times.log 2>&1
echo "$(tail -1 gpurun_out/${T}_layer_times.log) | $(awk '$1==3||$1==5||$1==15||$1==17||$1==23||$1==27||$1==39||$1==51{printf "%s:%s ", $1, $3}' gpurun_out/${T}_layer_times.log)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
