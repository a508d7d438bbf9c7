#!/bin/bash
# Round-2 closing evidence (one GPU): tests, bench lines, per-layer times, ncu launch list / traffic / full capture of conv_rs,
# training step, fused-TCN A/B, sanitizer.  Outputs under gpurun_out/r2f_*; copy what is kept into profiles/.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2f_pytest_gpu.log 2>&1; tail -3 $O/r2f_pytest_gpu.log
timeout 900 python bench.py > $O/r2f_bench_miso1_paper_bf16x3.json 2> $O/r2f_bench.err; cut -c1-300 $O/r2f_bench_miso1_paper_bf16x3.json; tail -2 $O/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_bench_reference_arm.json 2> $O/r2f_bench_ref.err; cut -c1-200 $O/r2f_bench_reference_arm.json
timeout 900 python bench.py --workload pipeline_paper --steps 3 --warmup 3 --no-cpu-baseline > $O/r2f_bench_pipeline_paper_bf16x3.json 2> $O/r2f_bench_pp.err; cut -c1-300 $O/r2f_bench_pipeline_paper_bf16x3.json; tail -2 $O/r2f_bench_pp.err
timeout 300 python tools/layer_times.py bf16x3 > $O/r2f_layer_times_bf16x3.log 2>&1; tail -1 $O/r2f_layer_times_bf16x3.log
MISO_TCN_FUSED=1 timeout 300 python bench.py --no-cpu-baseline > $O/r2f_bench_tcn_fused.json 2>/dev/null; cut -c1-260 $O/r2f_bench_tcn_fused.json
timeout 600 python tools/train_step.py --steps 5 --warmup 3 --cpu-baseline > $O/r2f_train_step_1gpu.json 2> $O/r2f_train.err; cut -c1-260 $O/r2f_train_step_1gpu.json; tail -2 $O/r2f_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2f.csv python tools/one_fwd.py bf16x3 2 > $O/ncu_launch.log 2>&1
python tools/condense_ncu.py launches $O/launches_r2f.csv $O/r2f_ncu_launches_miso1_paper_b16_bf16x3.csv "ncu --metrics gpu__time_duration.sum --clock-control none python tools/one_fwd.py bf16x3 2  (second forward; B=16 x 6 x 500 x 257, eager launches)" 2>/dev/null | head -12
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_rs --csv --log-file $O/traffic_r2f.csv python tools/one_fwd.py bf16x3 2 > $O/ncu_traffic.log 2>&1
python tools/condense_ncu.py traffic $O/traffic_r2f.csv $O/r2f_ncu_traffic_conv_rs_bf16x3.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_rs python tools/one_fwd.py bf16x3 2 (second forward)" | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rs_kernel -s 9 -c 1 -o $O/r2f_prof_conv_rs_9 -f python tools/one_fwd.py bf16x3 1 > $O/ncu_rs_9.log 2>&1
ncu -i $O/r2f_prof_conv_rs_9.ncu-rep --page raw --csv 2>/dev/null > $O/r2f_conv_rs_9_raw.csv
python tools/ncu_summary.py < $O/r2f_conv_rs_9_raw.csv > $O/r2f_ncu_full_conv_rs_bf16x3_launch9.csv; head -30 $O/r2f_ncu_full_conv_rs_bf16x3_launch9.csv | cut -c1-160
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_compute_sanitizer_memcheck_smoke.log 2>&1; tail -3 $O/r2f_compute_sanitizer_memcheck_smoke.log
rm -f $O/launches_r2f.csv $O/traffic_r2f.csv $O/r2f_conv_rs_9_raw.csv $O/r2f_prof_conv_rs_9.ncu-rep
