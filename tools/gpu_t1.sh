#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "istft or long_recording or dropin or stft" > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f.log; tail -25 gpurun_out/pytest_f.log
