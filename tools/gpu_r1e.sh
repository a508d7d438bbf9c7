#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for mode in bf16x3 bf16; do
  timeout 600 python bench.py --steps 5 --warmup 3 --conv-mode $mode --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$mode.json"))
print("$mode", round(d["value"]), "frames/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"]), {k[:12]:(round(v["ms_per_step"],2), round(v["tflops"],1)) for k,v in d["roofline"]["families"].items()}, d["clocks"])
PY
  tail -n 2 gpurun_out/bench_$mode.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bf16x3.csv \
    python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --conv-mode bf16x3 > gpurun_out/ncu_launch.log 2>&1
