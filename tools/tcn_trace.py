"""Phase stamps (clock64, CTA 0) of the fused TCN launch in the bench workload: per half-block, cycles relative to the half's start."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from misonet_b200 import _lib
from misonet_b200.model import MISO_1
wl = bench.WORKLOADS["miso1_paper"]
en, de = bench.LAYOUTS[wl["layout"]]
m = MISO_1(2, 6, len(en), list(en), list(de), "IN")
m.load_state_dict(bench.make_state_dict_np(m, 0))
m = m.cuda().eval(); m.conv_mode = "bf16x3"; m.use_graph = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else wl["B"]
x = bench.rand_spec(100, (B, 6, wl["T"], wl["F"]), "cuda")
lib = _lib.load()
buf = torch.zeros(32 * 16, dtype=torch.int64, device="cuda")
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    lib.miso_debug_tc_trace(buf.data_ptr(), -7, 0)
    m(x); torch.cuda.synchronize()
    lib.miso_debug_tc_trace(None, 0, 0)
h = buf.cpu().view(32, 16)
names = {0: "start", 1: "affine", 2: "dw done", 3: "gLN add", 4: "barrier2", 5: "vec", 6: "acc ready", 7: "epi done", 8: "sums", 10: "mma start", 11: "w full", 12: "a full", 13: "mma issued"}
for i in range(28):
    t0 = int(h[i, 0])
    if t0 == 0: continue
    print(i, " ".join(f"{names[k]}:{int(h[i, k]) - t0}" for k in sorted(names) if int(h[i, k])), "| next start", int(h[i + 1, 0]) - t0 if int(h[i + 1, 0]) else "-")
