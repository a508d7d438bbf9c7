#!/bin/bash
mkdir -p gpurun_out
for s in 9 4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rs_kernel -s $s -c 1 -o gpurun_out/prof_conv_rs_$s -f python tools/one_fwd.py bf16x3 1 > gpurun_out/ncu_rs_$s.log 2>&1
ncu -i gpurun_out/prof_conv_rs_$s.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/ncu_rs_$s.csv
cat gpurun_out/ncu_rs_$s.csv
done
