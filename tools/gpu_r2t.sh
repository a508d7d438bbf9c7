#!/bin/bash
# round 2, call T: coalesced channels-last output of conv_rs (in-place data gradients) -- tests, training step A/B
mkdir -p gpurun_out
T=${1:-r2t}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not fused_tcn" > gpurun_out/${T}_pytest_parity.log 2>&1; tail -3 gpurun_out/${T}_pytest_parity.log
MISO_DGRAD_RS_CL=1 timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -x > gpurun_out/${T}_pytest_training_cl.log 2>&1; tail -3 gpurun_out/${T}_pytest_training_cl.log
timeout 600 python tools/train_step.py --steps 5 --warmup 3 > gpurun_out/${T}_train_default.json 2> gpurun_out/${T}_train_default.err; cut -c1-200 gpurun_out/${T}_train_default.json; grep -o '"phases_ms_rank0": {[^}]*}' gpurun_out/${T}_train_default.json
MISO_DGRAD_RS_CL=1 timeout 600 python tools/train_step.py --steps 5 --warmup 3 > gpurun_out/${T}_train_cl.json 2> gpurun_out/${T}_train_cl.err; cut -c1-200 gpurun_out/${T}_train_cl.json; grep -o '"phases_ms_rank0": {[^}]*}' gpurun_out/${T}_train_cl.json; tail -2 gpurun_out/${T}_train_cl.err
