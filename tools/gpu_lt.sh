#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/layer_times.py bf16x3 > gpurun_out/lt_rs_bf16x3.log 2>&1; tail -1 gpurun_out/lt_rs_bf16x3.log
python - <<'PY'
from collections import defaultdict
d=defaultdict(lambda:[0,0.0,0.0])
for l in open('gpurun_out/lt_rs_bf16x3.log'):
    p=l.split()
    if len(p)>5 and p[0].isdigit():
        d[p[1]][0]+=1; d[p[1]][1]+=float(p[2]); d[p[1]][2]+=float(p[4])
for k,(n,us,gf) in sorted(d.items()): print(k, n, round(us/1e3,3),'ms', round(gf/us*1e-3*1e3,1) if us else 0,'TFLOP/s')
PY
