#!/bin/bash
# bench line + reference arm + ncu launch list + one full ncu capture of the dominant kernel
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --workload pipeline_ref --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pipeline.json 2> gpurun_out/bench_pipeline.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_fp32 -s 30 -c 3 -o gpurun_out/prof_conv_fp32 \
    python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_ours.json gpurun_out/bench_ref.json gpurun_out/bench_pipeline.json
tail -n 3 gpurun_out/bench_ours.err gpurun_out/bench_pipeline.err
