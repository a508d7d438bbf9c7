#!/bin/bash
# round 2, call E (2 GPUs): bucketed gradient all-reduce check, train step at 1 and 2 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
$TR --nproc-per-node 2 tools/ddp_check.py > gpurun_out/r2e_ddp_check_2gpu.log 2>&1; tail -4 gpurun_out/r2e_ddp_check_2gpu.log
python tools/train_step.py --steps 3 --warmup 2 > gpurun_out/r2e_train_step_1gpu.json 2> gpurun_out/r2e_train_1.err; cut -c1-600 gpurun_out/r2e_train_step_1gpu.json
$TR --nproc-per-node 2 tools/train_step.py --steps 3 --warmup 2 > gpurun_out/r2e_train_step_2gpu.json 2> gpurun_out/r2e_train_2.err; cut -c1-900 gpurun_out/r2e_train_step_2gpu.json; tail -3 gpurun_out/r2e_train_2.err
