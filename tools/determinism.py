"""Run-to-run reproducibility of the MISO_1 forward (eager and graph), per conv mode."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from misonet_b200 import synth
from misonet_b200.model import MISO_1
from oracle import weights, miso_net_torch as mnt
def rel(a, b): return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))
en, de = mnt.LAYOUTS["REF"]
cfg = mnt.NetConfig.miso1()
sd = weights.make_state_dict(cfg, 0)
m = MISO_1(2, 6, 7, list(en), list(de), "IN"); m.load_state_dict(sd); m = m.cuda().eval()
mix = torch.from_numpy(synth.random_spec(21, (2, 6, 24, 129))).cuda()
for mode in ("fp32", "bf16x3", "bf16"):
    m.conv_mode = mode
    for graph in (False, True):
        m.use_graph = graph
        outs = []
        taps = []
        with torch.no_grad():
            for _ in range(4):
                outs.append(m(mix).cpu().numpy())
                taps.append({n: m.tap(n, 2, 24, 129).cpu().numpy() for n in ("enc0", "enc1", "enc4", "tcn", "dec2", "dec5")})
        print(mode, "graph" if graph else "eager", "run-to-run:", [f"{rel(o, outs[0]):.1e}" for o in outs[1:]],
              "taps run1 vs run0:", {k: f"{rel(taps[1][k], taps[0][k]):.1e}" for k in taps[0]}, flush=True)
    # eager vs graph
