#!/bin/bash
# round 2, call F: tcgen05 weight gradient -- gradient parity tests, training-step time (all layers / DenseBlocks only / legacy)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -q -x > gpurun_out/r2f_pytest_training.log 2>&1; tail -3 gpurun_out/r2f_pytest_training.log
for v in 1 2 0 1 2; do
MISO_WGRAD_TC=$v timeout 600 python tools/train_step.py --steps 5 --warmup 3 > gpurun_out/r2f_train_step_1gpu_tc$v.json 2> gpurun_out/r2f_train_$v.err; python -c "import json;d=json.load(open('gpurun_out/r2f_train_step_1gpu_tc$v.json'));print('MISO_WGRAD_TC=$v',d['ms_per_step'],d['phases_ms_rank0'])"; tail -2 gpurun_out/r2f_train_$v.err
done
