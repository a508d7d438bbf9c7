"""Small forward (two frame tiles in the TCN) for compute-sanitizer / debugging: prints the error against the oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import _model
from conftest import rel_err
from misonet_b200 import synth
from oracle import miso_net_torch as mnt
layout = sys.argv[1] if len(sys.argv) > 1 else "REF"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
F = 129 if layout == "REF" else 257
m, cfg, sd = _model("miso1", 5, layout=layout)
m.use_graph = False
mix = synth.random_spec(7, (B, 6, T, F))
with torch.no_grad():
    y = m(torch.from_numpy(mix).cuda())
    torch.cuda.synchronize()
    tcn = m.tap("tcn", B, T, F).cpu().numpy()
y = y.cpu().numpy()
ref = mnt.miso1_forward(sd, cfg, torch.from_numpy(mix)).numpy()
print("REL_ERR", rel_err(y, ref), "tcn finite", bool(np.isfinite(tcn).all()))
