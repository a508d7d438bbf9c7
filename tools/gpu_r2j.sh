#!/bin/bash
# round 2, call J: conv_rs layer variants (first conv, transposed convs) -- parity tests, bench A/B, per-layer times
mkdir -p gpurun_out
MISO_TC_DEBUG=1 timeout 300 python -c "
import torch, sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from test_gpu_parity import _model
from conftest import rel_err
from misonet_b200 import synth
from oracle import miso_net_torch as mnt
m,cfg,sd=_model('miso1',5,layout='PAPER'); m.use_graph=False
mix=synth.random_spec(7,(2,6,40,257))
ref=mnt.miso1_forward(sd,cfg,torch.from_numpy(mix)).numpy()
with torch.no_grad(): y=m(torch.from_numpy(mix).cuda()).cpu().numpy()
print('REL_ERR', rel_err(y,ref))
" > gpurun_out/r2j_debug.log 2>&1; grep -c "conv_rs: kind" gpurun_out/r2j_debug.log; grep "conv_rs: kind [1-4]" gpurun_out/r2j_debug.log | cut -c1-200; grep REL_ERR gpurun_out/r2j_debug.log; tail -3 gpurun_out/r2j_debug.log | cut -c1-300
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2j_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2j_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_default.json 2> gpurun_out/r2j_bench_default.err; cut -c1-330 gpurun_out/r2j_bench_default.json; tail -2 gpurun_out/r2j_bench_default.err
MISO_RS_VARIANTS=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_novar.json 2> gpurun_out/r2j_bench_novar.err; cut -c1-330 gpurun_out/r2j_bench_novar.json
timeout 300 python tools/layer_times.py > gpurun_out/r2j_layer_times.log 2>&1; tail -1 gpurun_out/r2j_layer_times.log
