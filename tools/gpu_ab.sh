#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  MISO_TC_DENSE=$v timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_var$v.log 2>&1
  echo "DENSE=$v x3:"; sed -n 10,12p gpurun_out/lt_var$v.log; sed -n 34,35p gpurun_out/lt_var$v.log; sed -n 62,63p gpurun_out/lt_var$v.log; tail -1 gpurun_out/lt_var$v.log
  MISO_TC_DENSE=$v timeout 300 python tools/layer_times.py bf16 > gpurun_out/lt_var${v}_bf16.log 2>&1; echo "bf16:"; tail -1 gpurun_out/lt_var${v}_bf16.log
done
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
