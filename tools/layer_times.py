"""Per-conv-launch device times (event nodes in the replayed graph) of the default bench workload."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from misonet_b200 import _lib
from misonet_b200.model import MISO_1
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
wl = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "miso1_paper"]
B = int(sys.argv[3]) if len(sys.argv) > 3 else wl["B"]
en, de = bench.LAYOUTS[wl["layout"]]
m = MISO_1(2, 6, len(en), list(en), list(de), "IN")
m.load_state_dict(bench.make_state_dict_np(m, 0))
m = m.cuda().eval(); m.conv_mode = mode
x = bench.rand_spec(100, (B, 6, wl["T"], wl["F"]), "cuda")
lib = _lib.load()
with torch.no_grad():
    for _ in range(3): m(x)
    lib.miso_prof_enable(1)
    m(x); torch.cuda.synchronize()
    lib.miso_prof_collect(-1, None, None, None, None)
    m(x); torch.cuda.synchronize()
cap = 4096
ms = (ctypes.c_double * cap)(); fl = (ctypes.c_double * cap)(); fam = (ctypes.c_int * cap)()
n = lib.miso_prof_dump(ms, fl, fam, cap)
tot = 0.0
for i in range(n):
    tot += ms[i]
    print(f"{i:3d} fam{fam[i]} {ms[i]*1e3:9.1f} us  {fl[i]/1e9:8.2f} GFLOP  {fl[i]/max(ms[i],1e-9)/1e9:8.1f} TFLOP/s")
print("total", round(tot, 3), "ms over", n, "launches")
