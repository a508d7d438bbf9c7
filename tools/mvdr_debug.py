"""GPU debug: compare every MVDR intermediate with the oracle's parts."""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from misonet_b200 import beamforming, synth
from oracle import miso_np

def rel(a, b): return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))
def al(x, a=256): return (x + a - 1) // a * a

for (B, F, M, T) in [(1, 9, 6, 40), (2, 33, 6, 50)]:
    src, mix = synth.mvdr_case(1, B, F, M, T)
    ref, parts = miso_np.apply_beamforming(src, mix, return_parts=True)
    s_t = torch.from_numpy(src).permute(0, 2, 3, 1).contiguous().cuda().unsqueeze(0)
    m_t = torch.from_numpy(mix).permute(0, 2, 3, 1).contiguous().cuda()
    y, w = beamforming.mvdr(s_t, m_t, return_weights=True)
    torch.cuda.synchronize()
    ws = beamforming._ws_cache[m_t.device]
    ctas = ((F + 31) // 32) * B
    ts = max(1, min(8, (296 + ctas - 1) // ctas))
    NV = 2 * M * (M + 1)
    nb_partial = B * 1 * ts * NV * F * 4
    partial = ws[:nb_partial].view(torch.float32).view(B, ts, NV, F).cpu().numpy().astype(np.float64)
    steer = ws[al(nb_partial):al(nb_partial) + B * F * M * 16].view(torch.float64).view(B, F, M, 2).cpu().numpy()
    steer = steer[..., 0] + 1j * steer[..., 1]
    tot = partial.sum(axis=1) / T      # [B, NV, F]
    Ps = np.zeros((B, F, M, M), complex); Pn = np.zeros((B, F, M, M), complex)
    k = 0
    for i in range(M):
        for j in range(i, M):
            Ps[:, :, i, j] = tot[:, 4 * k] + 1j * tot[:, 4 * k + 1]
            Pn[:, :, i, j] = tot[:, 4 * k + 2] + 1j * tot[:, 4 * k + 3]
            if i != j:
                Ps[:, :, j, i] = np.conj(Ps[:, :, i, j]); Pn[:, :, j, i] = np.conj(Pn[:, :, i, j])
            k += 1
    print(f"case B{B} F{F} M{M} T{T} tsplit {ts}")
    print("  scm_s", rel(Ps, parts["scm_s"]), " scm_n", rel(Pn, parts["scm_n"]))
    # steering before phase correction
    v = parts["eigvec"]; d0 = miso_np.steering_normalise(v)
    print("  steer(before phase)", rel(steer, d0))
    print("  weights", rel(w[0].cpu().numpy(), parts["weights"]))
    print("  out", rel(y[0].cpu().numpy(), ref))
    dpc = miso_np.phase_correction(d0.astype(np.complex128))
    print("  oracle steering (phase corrected) vs parts", rel(dpc, parts["steering"]))
