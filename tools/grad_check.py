"""Per-parameter gradient error table of the training path against torch autograd over the CPU oracle.
usage: python tools/grad_check.py [REF|PAPER] [mode] [B] [T] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_training as tt  # noqa: E402
from misonet_b200 import synth  # noqa: E402

layout = sys.argv[1] if len(sys.argv) > 1 else "REF"
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
T = int(sys.argv[4]) if len(sys.argv) > 4 else 40
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 4
F = 129 if layout == "REF" else 257
alpha = float(sys.argv[6]) if len(sys.argv) > 6 else None
m, cfg, sd = tt._model(seed, layout, mode, alpha)
mix = torch.from_numpy(synth.random_spec(11, (B, 6, T, F)))
up = torch.from_numpy(synth.random_spec(12, (B, 2, T, F)))
out_ref, g_ref, _ = tt._oracle_grads(sd, cfg, mix, upstream=up, dtype=torch.float64)
out = m(mix.cuda())
print("forward rel err", tt.rel_err(out.detach().cpu().numpy(), out_ref.numpy()))
out.backward(up.cuda())
rows = []
for k, p in m.named_parameters():
    g, r = p.grad.cpu().double().numpy().ravel(), g_ref[k].numpy().ravel()
    rows.append((tt.rel_err(g, r), k, float(np.linalg.norm(r)), float(np.linalg.norm(g))))
for e, k, nr, ng in rows:
    print(f"{e:10.3e}  |ref|={nr:10.3e} |ours|={ng:10.3e}  {k}")
print("worst", max((r for r in rows if not r[1].endswith("net.2.net.2.beta")), key=lambda r: r[0]))
num = sum((e * nr) ** 2 for e, k, nr, ng in rows if nr > 1e-3)
den = sum(nr ** 2 for e, k, nr, ng in rows)
print("total %.3e" % ((num / den) ** 0.5))
