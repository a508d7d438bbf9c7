#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one_fwd.py <<'PY'
import os, sys
import torch
sys.path.insert(0, os.getcwd())
import bench
from misonet_b200.model import MISO_1
mode = sys.argv[1]
wl = bench.WORKLOADS["miso1_paper"]
en, de = bench.LAYOUTS[wl["layout"]]
m = MISO_1(2, 6, len(en), list(en), list(de), "IN")
m.load_state_dict(bench.make_state_dict_np(m, 0))
m = m.cuda().eval(); m.conv_mode = mode; m.use_graph = False
x = bench.rand_spec(100, (wl["B"], 6, wl["T"], wl["F"]), "cuda")
with torch.no_grad():
    for _ in range(2): m(x)
    torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16x3.csv python /tmp/one_fwd.py bf16x3 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
