#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python tools/train_step.py --steps 5 --warmup 3 > gpurun_out/r2h_train_cl_$i.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r2h_train_cl_$i.json'));print('cl',d['ms_per_step'],d['phases_ms_rank0'])"
MISO_DGRAD_RS_CL=0 timeout 600 python tools/train_step.py --steps 5 --warmup 3 > gpurun_out/r2h_train_planes_$i.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r2h_train_planes_$i.json'));print('planes',d['ms_per_step'],d['phases_ms_rank0'])"
done
