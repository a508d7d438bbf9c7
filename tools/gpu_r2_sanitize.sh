#!/bin/bash
# compute-sanitizer memcheck of the round-2 kernels: smoke() (forward incl. the conv_rs variants, MVDR, STFT/ISTFT, one training
# step with the tcgen05 weight gradient) and a small PAPER-layout training step (strided / transposed weight-gradient variants)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke > gpurun_out/r2_compute_sanitizer_memcheck_smoke.log 2>&1; tail -4 gpurun_out/r2_compute_sanitizer_memcheck_smoke.log
MISO_TRAIN_GRAPH=0 timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python tools/train_step.py --layout PAPER --batch 2 --frames 24 --steps 1 --warmup 0 --one-step > gpurun_out/r2_compute_sanitizer_memcheck_train.log 2>&1; tail -4 gpurun_out/r2_compute_sanitizer_memcheck_train.log
