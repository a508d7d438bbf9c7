#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r1_c_bench_miso1_paper_bf16x3.json 2> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/r1_c_bench_miso1_paper_bf16x3.json; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --conv-mode bf16 --no-cpu-baseline > gpurun_out/r1_c_bench_miso1_paper_bf16.json 2> gpurun_out/bench_bf16.err; cut -c1-200 gpurun_out/r1_c_bench_miso1_paper_bf16.json
timeout 600 python bench.py --workload miso1_ref --no-cpu-baseline > gpurun_out/r1_c_bench_miso1_ref_bf16x3.json 2> gpurun_out/bench_miso1_ref.err; cut -c1-200 gpurun_out/r1_c_bench_miso1_ref_bf16x3.json
timeout 900 python bench.py --workload pipeline_ref --steps 3 --no-cpu-baseline > gpurun_out/r1_c_bench_pipeline_ref_bf16x3.json 2> gpurun_out/bench_pipeline.err; cut -c1-200 gpurun_out/r1_c_bench_pipeline_ref_bf16x3.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_c_bench_reference_arm.json 2> gpurun_out/bench_reference.err; cut -c1-200 gpurun_out/r1_c_bench_reference_arm.json
