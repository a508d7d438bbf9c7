#!/bin/bash
# Regenerates the round-2 evidence under gpurun_out/ (copy what is wanted into profiles/).  One GPU.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py > $O/r2_bench_miso1_paper_bf16x3.json 2> $O/r2_bench_default.err; cut -c1-300 $O/r2_bench_miso1_paper_bf16x3.json; tail -2 $O/r2_bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_ref.err; cut -c1-200 $O/r2_bench_reference_arm.json
timeout 900 python bench.py --workload pipeline_paper --steps 3 --warmup 3 > $O/r2_bench_pipeline_paper_bf16x3.json 2> $O/r2_bench_pp.err; cut -c1-300 $O/r2_bench_pipeline_paper_bf16x3.json; tail -2 $O/r2_bench_pp.err
timeout 900 python bench.py --workload pipeline_ref --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_pipeline_ref_bf16x3.json 2> $O/r2_bench_pr.err; cut -c1-200 $O/r2_bench_pipeline_ref_bf16x3.json
timeout 300 python tools/layer_times.py bf16x3 > $O/r2_layer_times_bf16x3.log 2>&1; tail -1 $O/r2_layer_times_bf16x3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2.csv python tools/one_fwd.py bf16x3 2 > $O/ncu_launch.log 2>&1
python tools/condense_ncu.py launches $O/launches_r2.csv $O/r2_ncu_launches_miso1_paper_b16_bf16x3.csv "ncu --metrics gpu__time_duration.sum --clock-control none python tools/one_fwd.py bf16x3 2  (second forward; B=16 x 6 x 500 x 257, eager launches)" | head -12
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_rs --csv --log-file $O/traffic_r2.csv python tools/one_fwd.py bf16x3 2 > $O/ncu_traffic.log 2>&1
python tools/condense_ncu.py traffic $O/traffic_r2.csv $O/r2_ncu_traffic_conv_rs_bf16x3.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_rs python tools/one_fwd.py bf16x3 2 (second forward)" | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rs_kernel -s 9 -c 1 -o $O/r2_prof_conv_rs_9 -f python tools/one_fwd.py bf16x3 1 > $O/ncu_rs_9.log 2>&1
ncu -i $O/r2_prof_conv_rs_9.ncu-rep --page raw --csv 2>/dev/null > $O/r2_conv_rs_9_raw.csv
python tools/ncu_summary.py < $O/r2_conv_rs_9_raw.csv > $O/r2_ncu_full_conv_rs_bf16x3_launch9.csv; head -40 $O/r2_ncu_full_conv_rs_bf16x3_launch9.csv | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 8 -c 1 -o $O/r2_prof_wgrad_tc -f python tools/train_step.py --steps 1 --warmup 0 --one-step > $O/ncu_wtc.log 2>&1
ncu -i $O/r2_prof_wgrad_tc.ncu-rep --page raw --csv 2>/dev/null > $O/r2_wgrad_tc_raw.csv
python tools/ncu_summary.py < $O/r2_wgrad_tc_raw.csv > $O/r2_ncu_full_wgrad_tc_bf16x3.csv; head -40 $O/r2_ncu_full_wgrad_tc_bf16x3.csv | cut -c1-160
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_ncu_train_b8_raw.csv python tools/train_step.py --batch 8 --steps 1 --warmup 0 --one-step > $O/r2_train_ncu.log 2>&1
python tools/condense_ncu.py launches $O/r2_ncu_train_b8_raw.csv $O/r2_ncu_launches_train_paper_b8_bf16x3.csv "ncu --metrics gpu__time_duration.sum --clock-control none python tools/train_step.py --batch 8 --steps 1 --warmup 0 --one-step" | head -24
rm -f $O/launches_r2.csv $O/traffic_r2.csv $O/r2_ncu_train_b8_raw.csv
