#!/bin/bash
# Regenerates the r1_d (training path) evidence under profiles/ on a GPU box: run through gpurun from the repo root.
# 2-GPU lines need `gpurun --gpus 2`.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q > gpurun_out/pytest_training.log 2>&1; tail -3 gpurun_out/pytest_training.log
timeout 300 python tools/train_step.py --batch 8 --steps 5 --warmup 3 --cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r1_d_train_step_1gpu.json
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      tools/train_step.py --batch 8 --steps 5 --warmup 3 2>/dev/null | grep "^{" > gpurun_out/r1_d_train_step_2gpu.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
      tools/ddp_check.py 2>&1 | grep -v Warning | tail -4 > gpurun_out/r1_d_ddp_check_2gpu.log
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_train_b8_raw.csv \
    python tools/train_step.py --batch 8 --steps 1 --warmup 0 > gpurun_out/train_ncu.log 2>&1
python tools/condense_ncu.py launches gpurun_out/ncu_train_b8_raw.csv gpurun_out/r1_d_ncu_launches_train_paper_b8_bf16x3.csv \
    "ncu --metrics gpu__time_duration.sum --clock-control none python tools/train_step.py --batch 8 --steps 1 --warmup 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_taps -s 30 -c 1 -o gpurun_out/r1_d_wgrad_taps \
    python tools/train_step.py --batch 8 --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1
timeout 700 compute-sanitizer --tool memcheck --print-limit 5 python tools/train_step.py --layout PAPER --batch 2 --frames 24 \
    --steps 1 --warmup 0 > gpurun_out/r1_d_compute_sanitizer_memcheck_train.log 2>&1
