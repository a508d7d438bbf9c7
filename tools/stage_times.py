"""The small HBM-bound stages alone against their algorithmic bytes: STFT, ISTFT, the
alignment / uPIT pair reductions.  L2 is flushed between iterations; median of the CUDA-event times."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from misonet_b200 import audio, criterion

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T, F, M, S, N = 501, 129, 6, 2, 32000
peak = 6552.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


g = torch.Generator(device="cuda").manual_seed(0)
wav = torch.randn(B, N, M, device="cuda", generator=g) * 0.1
spec = audio.stft(wav)
est = torch.view_as_complex(torch.randn(B, S, T, F, 2, device="cuda", generator=g))
ref = torch.view_as_complex(torch.randn(B, S, T, F, 2, device="cuda", generator=g))
rows = []
for name, fn, nbytes in [
    ("stft [B,32000,6] -> [B,6,501,129]", lambda: audio.stft(wav), B * (M * N * 4 + M * T * F * 8)),
    ("istft [B,2,501,129] -> [B,2,32000]", lambda: audio.istft(est), B * S * (T * F * 8 + (T - 1) * 64 * 4)),
    ("alignment distance + argmin (tester.py:1043-1059)", lambda: criterion.pair_decide(est, ref, 0), B * 2 * S * T * F * 8),
    ("uPIT L1 triple + argmin + mean (criterion.py:8-63)", lambda: criterion.pair_decide(est, ref, 1, want_loss=True), B * 2 * S * T * F * 8),
]:
    ms = timed(fn)
    rows.append({"stage": name, "B": B, "ms_median": ms, "algorithmic_MB": nbytes / 1e6, "achieved_GBs": nbytes / ms / 1e6,
                 "frac_of_hbm_peak": nbytes / ms / 1e6 / peak})
print(json.dumps({"peak_GBs": peak, "l2": "512 MiB buffer written between iterations", "stages": rows}, indent=1))
