#!/bin/bash
# round 2, call F: tcgen05 weight gradient -- gradient parity tests, training-step time
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -q -x > gpurun_out/r2f_pytest_training.log 2>&1; tail -30 gpurun_out/r2f_pytest_training.log
timeout 600 python tools/train_step.py --steps 3 --warmup 2 > gpurun_out/r2f_train_step_1gpu.json 2> gpurun_out/r2f_train_1.err; cut -c1-700 gpurun_out/r2f_train_step_1gpu.json; tail -3 gpurun_out/r2f_train_1.err
MISO_WGRAD_TC=0 timeout 600 python tools/train_step.py --steps 3 --warmup 2 > gpurun_out/r2f_train_step_1gpu_legacy.json 2> gpurun_out/r2f_train_1l.err; cut -c1-500 gpurun_out/r2f_train_step_1gpu_legacy.json
