#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r1_c_compute_sanitizer_memcheck_smoke.log 2>&1; echo "sanitizer rc=$?"; grep -v "^make\|Nothing to be done" gpurun_out/r1_c_compute_sanitizer_memcheck_smoke.log | tail -8
