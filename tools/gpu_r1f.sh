#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
MISO_TC_DEBUG=1 timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/layer_times_bf16x3.log 2> gpurun_out/geom_bf16x3.log; tail -2 gpurun_out/layer_times_bf16x3.log
timeout 300 python tools/layer_times.py bf16 > gpurun_out/layer_times_bf16.log 2>&1; tail -2 gpurun_out/layer_times_bf16.log
for mode in bf16x3 bf16; do
  timeout 600 python bench.py --steps 10 --warmup 3 --conv-mode $mode --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$mode.json"))
print("$mode", round(d["value"]), "frames/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"]), {k[:12]:(round(v["ms_per_step"],2), round(v["tflops"],1)) for k,v in d["roofline"]["families"].items()}, d["clocks"])
PY
  tail -n 2 gpurun_out/bench_$mode.err
done
