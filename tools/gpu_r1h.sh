#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_tc --csv --log-file gpurun_out/traffic_bf16x3.csv python tools/one_fwd.py bf16x3 2 > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json
