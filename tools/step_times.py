"""Per-step device times of the default bench workload + nvidia-smi clocks/power at 20 ms, to see throttling."""
import os, subprocess, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from misonet_b200.model import MISO_1
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wl = bench.WORKLOADS["miso1_paper"]
en, de = bench.LAYOUTS[wl["layout"]]
m = MISO_1(2, 6, len(en), list(en), list(de), "IN")
m.load_state_dict(bench.make_state_dict_np(m, 0))
m = m.cuda().eval(); m.conv_mode = mode
x = bench.rand_spec(100, (wl["B"], 6, wl["T"], wl["F"]), "cuda")
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
with torch.no_grad():
    for _ in range(3): m(x)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
    ev[0].record()
    for i in range(nsteps):
        m(x); ev[i + 1].record()
    torch.cuda.synchronize()
time.sleep(0.1); smi.terminate()
out = smi.stdout.read().strip().splitlines()
print(mode, "step ms:", [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(nsteps)])
print("smi samples (MHz, W, powercap, reasons):", out[::3])
