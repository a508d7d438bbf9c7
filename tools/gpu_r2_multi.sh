#!/bin/bash
# BASELINE configs[3] (training step, 8 utterances per GPU) and configs[4] (8000-frame recording, 16 chunks) at the rank
# counts given as arguments (default: 1 2).  Run under gpurun --gpus N with N = the largest count.
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in ${@:-1 2}; do
  if [ "$n" = 1 ]; then
    timeout 600 python tools/train_step.py --steps 5 --warmup 3 2> $O/r2_train_${n}.err | grep "^{" > $O/r2_train_step_${n}gpu.json
    timeout 600 python tools/long_recording.py 3 2> $O/r2_long_${n}.err | grep "^{" > $O/r2_long_recording_${n}gpu.json
  else
    timeout 600 $TR --master-port 2951$n --nproc-per-node $n tools/train_step.py --steps 5 --warmup 3 2> $O/r2_train_${n}.err | grep "^{" > $O/r2_train_step_${n}gpu.json
    timeout 600 $TR --master-port 2952$n --nproc-per-node $n tools/long_recording.py 3 2> $O/r2_long_${n}.err | grep "^{" > $O/r2_long_recording_${n}gpu.json
  fi
  python - <<PY
import json
d=json.load(open("$O/r2_train_step_${n}gpu.json")); print("train n=$n", round(d["value"]), "frames/s", round(d["ms_per_step"],2), "ms", d["collective"]["ms_per_step_without_allreduce"], d["collective"]["exposed_ms"])
d=json.load(open("$O/r2_long_recording_${n}gpu.json")); print("long  n=$n", round(d["frames_per_s"]), "frames/s", round(d["ms_per_recording"],2), "ms", d["bit_identical_to_one_rank"], d["rel_diff_to_one_rank"])
PY
done
