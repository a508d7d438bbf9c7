#!/bin/bash
# round 2, call I: full GPU suite + training step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2i_pytest_gpu.log
timeout 600 python tools/train_step.py --steps 5 --warmup 3 --cpu-baseline > gpurun_out/r2i_train_step_1gpu.json 2> gpurun_out/r2i_train.err; python -c "import json;d=json.load(open('gpurun_out/r2i_train_step_1gpu.json'));print(d['ms_per_step'],d['phases_ms_rank0'],d['cpu_baseline'])"; tail -2 gpurun_out/r2i_train.err
