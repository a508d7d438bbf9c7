#!/bin/bash
# ncu full captures of the dominant kernel + launch list + pipeline workload + reference arm
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16x3.csv python tools/one_fwd.py bf16x3 2 > gpurun_out/ncu_launch.log 2>&1
for spec in "11 l11_cin160_f127" "97 l97_cin144_f255"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $1 -c 1 -f -o gpurun_out/prof_conv_tc_$2 python tools/one_fwd.py bf16x3 1 > gpurun_out/ncu_full_$2.log 2>&1
  tail -2 gpurun_out/ncu_full_$2.log
done
timeout 600 python bench.py --workload pipeline_ref --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pipeline_bf16x3.json 2> gpurun_out/bench_pipeline.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_pipeline_bf16x3.json"))
print("pipeline", round(d["value"]), "frames/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"]), {k[:12]:(round(v["ms_per_step"],2), round(v["tflops"],1)) for k,v in d["roofline"]["families"].items()})
PY
tail -n 3 gpurun_out/bench_pipeline.err
timeout 600 python bench.py --workload miso1_ref --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_miso1_ref_bf16x3.json 2> gpurun_out/bench_miso1_ref.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_miso1_ref_bf16x3.json"))
print("miso1_ref", round(d["value"]), "frames/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"]))
PY
