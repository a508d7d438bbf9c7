#!/bin/bash
# round-1c evidence: launch list, DRAM traffic, full captures, layer times, bench lines, MVDR stage
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/one_fwd.py bf16x3 2 > gpurun_out/ncu_launch.log 2>&1
python tools/condense_ncu.py launches gpurun_out/launches_r1c.csv gpurun_out/r1_c_ncu_launches_miso1_paper_b16_bf16x3.csv "ncu --metrics gpu__time_duration.sum --clock-control none python tools/one_fwd.py bf16x3 2  (second forward; B=16 x 6 x 500 x 257, eager launches)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_rs --csv --log-file gpurun_out/traffic_r1c.csv python tools/one_fwd.py bf16x3 2 > gpurun_out/ncu_traffic.log 2>&1
python tools/condense_ncu.py traffic gpurun_out/traffic_r1c.csv gpurun_out/r1_c_ncu_traffic_conv_rs_bf16x3.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_rs python tools/one_fwd.py bf16x3 2 (second forward)"
for s in 9 14; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rs_kernel -s $s -c 1 -o gpurun_out/prof_conv_rs_$s -f python tools/one_fwd.py bf16x3 1 > gpurun_out/ncu_rs_$s.log 2>&1
ncu -i gpurun_out/prof_conv_rs_$s.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r1_c_ncu_full_conv_rs_bf16x3_launch$s.csv
head -12 gpurun_out/r1_c_ncu_full_conv_rs_bf16x3_launch$s.csv | cut -c1-200
done
timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/r1_c_layer_times_bf16x3.log 2>&1; tail -1 gpurun_out/r1_c_layer_times_bf16x3.log
timeout 300 python tools/layer_times.py bf16 > gpurun_out/r1_c_layer_times_bf16.log 2>&1; tail -1 gpurun_out/r1_c_layer_times_bf16.log
timeout 300 python tools/mvdr_times.py 32 500 257 > gpurun_out/r1_c_mvdr_stage_paper.json 2>/dev/null; cat gpurun_out/r1_c_mvdr_stage_paper.json
timeout 300 python tools/mvdr_times.py 32 501 129 > gpurun_out/r1_c_mvdr_stage_ref.json 2>/dev/null
