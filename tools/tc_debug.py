"""GPU debug: tcgen05 conv modes vs the fp32 path, tap by tap."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from misonet_b200 import synth
from misonet_b200.model import MISO_1
from oracle import weights, miso_net_torch as mnt

def rel(a, b): return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))
layout = sys.argv[1] if len(sys.argv) > 1 else "REF"
en, de = mnt.LAYOUTS[layout]
cfg = mnt.NetConfig.miso1(layout=layout)
sd = weights.make_state_dict(cfg, 0)
m = MISO_1(2, 6, len(en), list(en), list(de), "IN"); m.load_state_dict(sd); m = m.cuda().eval()
F = 129 if layout == "REF" else 257
for (B, T) in [(2, 20), (1, 70)]:
    mix = torch.from_numpy(synth.random_spec(7 + B, (B, 6, T, F))).cuda()
    res = {}
    for mode in ("fp32", "bf16x3", "bf16"):
        m.conv_mode = mode
        with torch.no_grad():
            y = m(mix)
        torch.cuda.synchronize()
        taps = {n: m.tap(n, B, T, F).cpu().numpy() for n in ["enc0", "enc1", "enc2", "enc4", "enc5", "tcn", "dec0", "dec2", "dec4", "dec5"]}
        taps["y"] = y.cpu().numpy()
        res[mode] = taps
    for mode in ("bf16x3", "bf16"):
        print(layout, B, T, mode, {k: f"{rel(res[mode][k], res['fp32'][k]):.2e}" for k in res[mode]}, flush=True)
    ref = mnt.miso1_forward(sd, cfg, mix.cpu()).numpy()
    print("   vs oracle:", {mode: f"{rel(res[mode]['y'], ref):.2e}" for mode in res}, flush=True)
