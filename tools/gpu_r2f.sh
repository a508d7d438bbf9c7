#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -q > gpurun_out/r2f_pytest_training.log 2>&1; tail -4 gpurun_out/r2f_pytest_training.log
timeout 600 python tools/train_step.py --steps 5 --warmup 3 > gpurun_out/r2f_train_step_1gpu.json 2> gpurun_out/r2f_train.err; python -c "import json;d=json.load(open('gpurun_out/r2f_train_step_1gpu.json'));print(d['ms_per_step'],d['phases_ms_rank0'])"; tail -2 gpurun_out/r2f_train.err
