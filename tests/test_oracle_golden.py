"""Pin the CPU oracle (oracle/) against outputs of the real reference
(tests/golden/*.npz, produced by oracle/make_golden.py in the build container)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import miso_np, weights
from oracle import miso_net_torch as mnt
from misonet_b200 import synth


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_stft_matches_reference():
    g = _load("stft_ref.npz")
    for i, (nperseg, noverlap) in enumerate(g["params"]):
        y = miso_np.stft(g[f"x{i}"], int(nperseg), int(noverlap))
        assert y.shape == g[f"y{i}"].shape
        assert y.shape[1] == miso_np.stft_num_frames(g[f"x{i}"].shape[0], int(nperseg), int(noverlap))
        assert rel_err(y, g[f"y{i}"]) < 2e-6


def test_istft_matches_reference_and_inverts_stft():
    g = _load("istft_ref.npz")
    for i, (nperseg, noverlap) in enumerate(g["params"]):
        y = miso_np.istft(g[f"spec{i}"], int(nperseg), int(noverlap))
        assert y.shape == g[f"wav{i}"].shape
        assert rel_err(y, g[f"wav{i}"]) < 2e-6
    # round trip: istft(stft(x)) == x wherever whole hops fit (scipy pads the tail with zeros)
    rng = np.random.default_rng(5)
    x = (0.1 * rng.standard_normal((64 * 30, 2))).astype(np.float32)
    spec = miso_np.stft(x, 256, 192)                    # [M, T, F]
    for m in range(2):
        back = miso_np.istft(spec[m], 256, 192)
        assert rel_err(back[:x.shape[0]], x[:, m]) < 1e-6


def test_mvdr_matches_reference():
    g = _load("mvdr_ref.npz")
    for i in range(2):
        y = miso_np.apply_beamforming(g[f"src{i}"], g[f"mix{i}"])
        assert y.dtype == np.complex64 and y.shape == g[f"y{i}"].shape
        assert rel_err(y, g[f"y{i}"]) < 1e-5


def test_mvdr_phase_scan_equivalence():
    """The sequential recurrence (tester.py:1163-1166) equals a prefix product of unit
    phasors on the uncorrected vectors -- the form the CUDA kernel uses."""
    rng = np.random.default_rng(0)
    d = rng.standard_normal((2, 40, 6)) + 1j * rng.standard_normal((2, 40, 6))
    seq = miso_np.phase_correction(d)
    c = np.sum(d[:, 1:] * d[:, :-1].conj(), axis=-1)
    ph = np.concatenate([np.ones((2, 1), complex), np.cumprod(np.exp(-1j * np.angle(c)), axis=1)], axis=1)
    assert rel_err(d * ph[..., None], seq) < 1e-12


def test_losses_match_reference():
    g = _load("loss_ref.npz")
    for i in range(2):
        loss, idx, _ = miso_np.loss_upit(g[f"est{i}"], g[f"ref{i}"])
        assert abs(loss - g[f"upit{i}"]) <= 2e-6 * abs(g[f"upit{i}"])
        assert idx[0] == 1          # utterance 0 was built to prefer the swapped order
        le = miso_np.loss_enhance(g[f"e1_{i}"], g[f"r1_{i}"])
        assert abs(le - g[f"enh{i}"]) <= 2e-6 * abs(g[f"enh{i}"])


@pytest.mark.parametrize("kind", ["miso1", "miso3"])
def test_net_matches_reference(kind):
    g = _load(f"net_ref_{kind}.npz")
    cfg = mnt.NetConfig.miso1() if kind == "miso1" else mnt.NetConfig.miso3()
    sd = weights.make_state_dict(cfg, 0 if kind == "miso1" else 1)
    assert weights.state_dict_digest(sd) == str(g["weights_digest"])
    for b, t in ((2, 20), (1, 11)):
        mix = torch.from_numpy(synth.random_spec(7 + b, (b, 6, t, 129)))
        if kind == "miso1":
            x = torch.cat((mix.real, mix.imag), dim=1)
        else:
            a2 = torch.from_numpy(synth.random_spec(17 + b, (b, 1, t, 129)))
            a3 = torch.from_numpy(synth.random_spec(27 + b, (b, 1, t, 129)))
            x = torch.cat((mix.real, a2.real, a3.real, mix.imag, a2.imag, a3.imag), dim=1)
        with torch.no_grad():
            y, taps = mnt.net_forward(sd, cfg, x, return_taps=True)
        yc = mnt._to_complex(y).numpy()
        assert rel_err(yc, g[f"y_b{b}"]) < 5e-6
        assert rel_err(taps["enc0"].numpy()[:, :, ::3, ::9], g[f"enc0_sub_b{b}"]) < 5e-6
        for name in ("enc4", "enc6", "dec0", "dec2"):
            assert rel_err(taps[name].numpy(), g[f"{name}_b{b}"]) < 5e-6, name
        assert rel_err(taps["tcn"].numpy(), g[f"tcn_b{b}"]) < 5e-6


def test_miso1_inference_matches_reference():
    g = _load("miso1_inference_ref.npz")
    cfg = mnt.NetConfig.miso1()
    sd = weights.make_state_dict(cfg, 0)
    for i in range(2):
        out, _ = mnt.miso1_inference(sd, cfg, torch.from_numpy(g[f"mix{i}"]), int(g[f"ref_ch{i}"]))
        for k in range(2):
            assert rel_err(out[k].numpy(), g[f"spk{k}_{i}"]) < 5e-6


def test_paper_layout_shapes():
    """PAPER layout (8 blocks, 257 bins, TCN width 384) runs and reduces F to 1;
    parameter count matches SURVEY.md section 8(c) (8,579,704)."""
    cfg = mnt.NetConfig.miso1(layout="PAPER")
    shapes = weights.param_shapes(cfg)
    assert sum(int(np.prod(s)) for s in shapes.values()) == 8579704
    sd = weights.make_state_dict(cfg, 3)
    mix = torch.from_numpy(synth.random_spec(1, (1, 6, 6, 257)))
    y = mnt.miso1_forward(sd, cfg, mix)
    assert y.shape == (1, 2, 6, 257) and y.dtype == torch.complex64
    assert torch.isfinite(torch.view_as_real(y)).all()


def test_ref_param_count():
    assert sum(int(np.prod(s)) for s in weights.param_shapes(mnt.NetConfig.miso1()).values()) == 2587384
    assert sum(int(np.prod(s)) for s in weights.param_shapes(mnt.NetConfig.miso3()).values()) == 2587382


def _training_case(kind, g):
    """Inputs of oracle/make_golden.py:golden_training and the oracle's loss + autograd gradients for them."""
    cfg = mnt.NetConfig.miso1() if kind == "miso1" else mnt.NetConfig.miso3()
    sd = weights.make_state_dict(cfg, 0 if kind == "miso1" else 1)
    for k in sd:
        if k.startswith("TCN.") and k.endswith(".net.1.weight"):
            sd[k] = torch.ones_like(sd[k])
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    b, t = 2, 12
    mix = torch.from_numpy(synth.random_spec(61, (b, 6, t, 129)))
    if kind == "miso1":
        refs = torch.from_numpy(synth.random_spec(62, (b, 2, t, 129)))
        y = mnt.net_forward(sdr, cfg, torch.cat((mix.real, mix.imag), dim=1))
        est = torch.complex(y[:, :2], y[:, 2:])
        e, r = est.unsqueeze(2), refs.unsqueeze(1)
        pair = ((e.real - r.real).abs().sum((3, 4)) + (e.imag - r.imag).abs().sum((3, 4)) +
                (torch.sqrt(e.real ** 2 + e.imag ** 2 + 1e-8) - r.abs()).abs().sum((3, 4)))            # criterion.py:27-33
        per = torch.stack([pair[:, 0, 0] + pair[:, 1, 1], pair[:, 0, 1] + pair[:, 1, 0]], dim=1)
        loss = per.min(dim=1).values.mean()
    else:
        a2 = torch.from_numpy(synth.random_spec(63, (b, 1, t, 129)))
        a3 = torch.from_numpy(synth.random_spec(64, (b, 1, t, 129)))
        refs = torch.from_numpy(synth.random_spec(65, (b, 1, t, 129)))
        y = mnt.net_forward(sdr, cfg, torch.cat((mix.real, a2.real, a3.real, mix.imag, a2.imag, a3.imag), dim=1))
        est = torch.complex(y[:, :1], y[:, 1:])
        loss = ((est.real - refs.real).abs().sum() + (est.imag - refs.imag).abs().sum() +
                (torch.sqrt(est.real ** 2 + est.imag ** 2 + 1e-8) - refs.abs()).abs().sum()) / b       # criterion.py:131-139
    loss.backward()
    return est.detach(), loss.detach(), {k: v.grad for k, v in sdr.items()}


@pytest.mark.parametrize("kind", ["miso1", "miso3"])
def test_training_step_matches_reference(kind):
    """The oracle's training step (network + loss + autograd) against the REAL reference's model(mix) -> loss ->
    loss.backward() (tests/golden/train_ref.npz): this is what pins the gradient parity tests of the CUDA path."""
    g = _load("train_ref.npz")
    est, loss, grads = _training_case(kind, g)
    assert rel_err(est.numpy(), g[f"{kind}_est"]) < 5e-6
    assert abs(float(loss) - float(g[f"{kind}_loss"])) <= 2e-6 * abs(float(g[f"{kind}_loss"]))
    norms = np.array([float(v.norm()) for v in grads.values()])
    ref_norms = g[f"{kind}_grad_norms"]
    scale = ref_norms.max()
    big = ref_norms > 1e-6 * scale                      # the gLN betas ahead of an InstanceNorm1d have zero gradient
    assert np.all(np.abs(norms[big] - ref_norms[big]) <= 2e-4 * ref_norms[big])          # measured 2.5e-5
    assert np.all(norms[~big] <= 1e-5 * scale)
    n = 0
    for key in g.files:
        if key.startswith(f"{kind}_grad::"):
            k = key.split("::", 1)[1]
            assert rel_err(grads[k].numpy(), g[key]) < 5e-5, k                # measured 2e-6
            n += 1
    assert n >= 4
