#!/bin/bash
# round 2, call G: training-step launch list (ncu) with the tcgen05 weight gradient
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_ncu_train_b8_raw.csv \
    python tools/train_step.py --batch 8 --steps 1 --warmup 0 > gpurun_out/r2g_train_ncu.log 2>&1
python tools/condense_ncu.py launches gpurun_out/r2g_ncu_train_b8_raw.csv gpurun_out/r2_ncu_launches_train_paper_b8_bf16x3.csv \
    "ncu --metrics gpu__time_duration.sum --clock-control none python tools/train_step.py --batch 8 --steps 1 --warmup 0"
tail -2 gpurun_out/r2g_train_ncu.log
