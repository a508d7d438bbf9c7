"""One eager (no CUDA graph) MISO_1 forward of the bench workload: the target of ncu captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from misonet_b200.model import MISO_1
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
nfwd = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = bench.WORKLOADS["miso1_paper"]
en, de = bench.LAYOUTS[wl["layout"]]
m = MISO_1(2, 6, len(en), list(en), list(de), "IN")
m.load_state_dict(bench.make_state_dict_np(m, 0))
m = m.cuda().eval(); m.conv_mode = mode; m.use_graph = False
x = bench.rand_spec(100, (wl["B"], 6, wl["T"], wl["F"]), "cuda")
with torch.no_grad():
    for _ in range(nfwd): m(x)
    torch.cuda.synchronize()
