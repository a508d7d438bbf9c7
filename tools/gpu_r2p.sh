#!/bin/bash
# round 2, call P: fused TCN iteration -- quick parity, the fused launch's time, bench
mkdir -p gpurun_out
T=${1:-r2p}
timeout 200 python tools/tcn_dbg.py PAPER 300 2 2>&1 | tail -2
timeout 300 python tools/layer_times.py > gpurun_out/${T}_layer_times.log 2>&1; grep fam2 gpurun_out/${T}_layer_times.log | head -3; tail -1 gpurun_out/${T}_layer_times.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-330 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
