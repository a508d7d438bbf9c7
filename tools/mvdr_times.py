"""MVDR stage alone (miso_mvdr_fwd: SCM -> eigenvector -> solve -> apply) against its HBM roofline.
Algorithmic bytes: (M + S*M) * T*F*8 read + S * T*F*8 written per utterance."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from misonet_b200.beamforming import mvdr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 501
F = int(sys.argv[3]) if len(sys.argv) > 3 else 129
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
S, M = 2, 6
g = torch.Generator(device="cuda").manual_seed(1)
mix = torch.view_as_complex(torch.randn(B, M, T, F, 2, device="cuda", generator=g))
# rank-1-ish sources (iid Gaussians make the eigenvector ill conditioned)
steer = torch.view_as_complex(torch.randn(S, B, M, 1, F, 2, device="cuda", generator=g))
sig = torch.view_as_complex(torch.randn(S, B, 1, T, F, 2, device="cuda", generator=g))
src = steer * sig + 0.05 * torch.view_as_complex(torch.randn(S, B, M, T, F, 2, device="cuda", generator=g))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): mvdr(src, mix)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); mvdr(src, mix); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
alg = B * ((M + S * M) * T * F * 8 + S * T * F * 8)
peak = 6552.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"stage": "mvdr", "B": B, "T": T, "F": F, "S": S, "M": M, "ms_median": ms, "algorithmic_MB": alg / 1e6,
                  "achieved_GBs": alg / ms / 1e6, "peak_GBs": peak, "frac": alg / ms / 1e6 / peak,
                  "frames_per_s": B * T / ms * 1e3, "l2": "512 MiB buffer written between iterations"}))
