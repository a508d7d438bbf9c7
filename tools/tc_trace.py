"""Cycle-level event log of CTA 0 of one tensor-core conv launch (selected by cin, Fin) in the bench workload."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from misonet_b200 import _lib
from misonet_b200.model import MISO_1
mode, cin, fin = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
wl = bench.WORKLOADS["miso1_paper"]
en, de = bench.LAYOUTS[wl["layout"]]
m = MISO_1(2, 6, len(en), list(en), list(de), "IN")
m.load_state_dict(bench.make_state_dict_np(m, 0))
m = m.cuda().eval(); m.conv_mode = mode; m.use_graph = False
x = bench.rand_spec(100, (wl["B"], 6, wl["T"], wl["F"]), "cuda")
lib = _lib.load()
buf = torch.zeros(4 * 4096, dtype=torch.int64, device="cuda")
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    lib.miso_debug_tc_trace(buf.data_ptr(), cin, fin)
    m(x); torch.cuda.synchronize()
    lib.miso_debug_tc_trace(None, 0, 0)
h = buf.cpu()[:3 * 4096].view(3, 2048, 2)
dur = buf.cpu()[3 * 4096:3 * 4096 + 148].tolist()
if any(dur):
    print('per-CTA cycles: min', min(dur), 'max', max(dur), 'mean', sum(dur) / len(dur))
    print('   ', ' '.join(str(d // 1000) for d in dur))
t0 = min(int(h[r, 0, 1]) for r in range(3) if int(h[r, 0, 0]) != 0)
names = ["producer", "mma", "epilogue"]
for r in range(3):
    ev = [(int(h[r, i, 0]), int(h[r, i, 1]) - t0) for i in range(2048) if int(h[r, i, 0]) != 0]
    print(names[r], len(ev), "events")
    print("   ", " ".join(f"{tag}@{clk}" for tag, clk in ev[:120]))
