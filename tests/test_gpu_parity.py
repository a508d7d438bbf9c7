"""GPU parity tests: the CUDA path (through the C ABI, via the host mirror in misonet_b200/)
against the reference-generated golden fixtures (tests/golden/*.npz) and against the CPU oracle
(oracle/) on seeded inputs.

Tolerances.  north_star asks for <= 1e-3 relative error on complex spectrogram values and
bit-exact permutation / alignment decisions.  The fp32 path is held to much tighter bounds
here (written next to each assert) so that regressions show up long before 1e-3.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

REQUIRED_TOL = 1e-3          # north_star
EN = [24, 32, 32, 32, 32, 64, 128]
DE = [128, 64, 32, 32, 32, 32, 24]


def _g(name):
    return np.load(os.path.join(GOLDEN, name))


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _model(kind, seed, layout="REF"):
    from misonet_b200.model import MISO_1, MISO_3
    from oracle import weights
    from oracle import miso_net_torch as mnt
    en, de = mnt.LAYOUTS[layout]
    if kind == "miso1":
        cfg = mnt.NetConfig.miso1(layout=layout)
        m = MISO_1(2, 6, len(en), list(en), list(de), "IN")
    else:
        cfg = mnt.NetConfig.miso3(layout=layout)
        m = MISO_3(1, 6, len(en), list(en), list(de), "IN")
    sd = weights.make_state_dict(cfg, seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), cfg, sd


# ------------------------------------------------------------------------------- S1 STFT
def test_stft_golden():
    from misonet_b200 import audio
    g = _g("stft_ref.npz")
    for i, (nperseg, noverlap) in enumerate(g["params"]):
        x = torch.from_numpy(g[f"x{i}"]).to(_dev())
        y = audio.stft(x, int(nperseg), int(noverlap)).cpu().numpy()
        assert y.shape == g[f"y{i}"].shape and y.dtype == np.complex64
        assert rel_err(y, g[f"y{i}"]) < 5e-6          # fp32 radix-2 FFT vs pocketfft


def test_istft_golden_and_round_trip():
    """ISTFT (tester.py:979-990 applied to spec * scale) against the reference-generated fixture, against the oracle
    at the full chunk size, and as the inverse of the STFT kernel."""
    from misonet_b200 import audio
    from oracle import miso_np
    g = _g("istft_ref.npz")
    for i, (nperseg, noverlap) in enumerate(g["params"]):
        y = audio.istft(torch.from_numpy(g[f"spec{i}"]).cuda(), int(nperseg), int(noverlap)).cpu().numpy()
        assert y.shape == g[f"wav{i}"].shape
        assert rel_err(y, g[f"wav{i}"]) < 2e-6
    rng = np.random.default_rng(9)
    spec = (rng.standard_normal((3, 2, 501, 129)) + 1j * rng.standard_normal((3, 2, 501, 129))).astype(np.complex64)
    y = audio.istft(torch.from_numpy(spec).cuda()).cpu().numpy()
    assert y.shape == (3, 2, 32000)
    for b, s in [(0, 0), (2, 1)]:
        assert rel_err(y[b, s], miso_np.istft(spec[b, s])) < 2e-6
    x = torch.from_numpy((0.1 * rng.standard_normal((2, 32000, 6))).astype(np.float32)).cuda()
    back = audio.istft(audio.stft(x))                      # [B, M, T, F] -> [B, M, 32000]
    assert rel_err(back.permute(0, 2, 1).cpu().numpy(), x.cpu().numpy()) < 2e-6


def test_wave_to_int16_matches_numpy_astype():
    """tester.py:155-157: wave * MaxINT16 then astype(np.int16) (truncation toward zero) -- bit-exact."""
    from misonet_b200 import audio
    rng = np.random.default_rng(4)
    x = np.clip(0.3 * rng.standard_normal(100003), -0.9999, 0.9999).astype(np.float32)     # in range: numpy's overflow is undefined
    x[:6] = [0.0, 1.0 / 32767, -1.0 / 32767, 0.99999, -0.99999, 0.5]
    want = (x.astype(np.float64) * np.iinfo(np.int16).max).astype(np.int16)
    got = audio.to_int16(torch.from_numpy(x).cuda()).cpu().numpy()
    assert got.dtype == np.int16 and np.array_equal(got, want)


def test_stft_oracle_batched_full_size():
    from misonet_b200 import audio, synth
    from oracle import miso_np
    mix, _ = synth.make_utterance(0, n_samples=32000)
    mix2, _ = synth.make_utterance(1, n_samples=32000)
    x = torch.from_numpy(np.stack([mix, mix2])).to(_dev())           # [2, N, 6]
    y = audio.stft(x).cpu().numpy()
    assert y.shape == (2, 6, 501, 129)
    for b, m in enumerate((mix, mix2)):
        assert rel_err(y[b], miso_np.stft(m)) < 5e-6
    # ragged length (padded=True branch) and the 512-point transform
    x3 = torch.from_numpy(mix[:12345]).to(_dev())
    assert rel_err(audio.stft(x3, 512, 384).cpu().numpy(), miso_np.stft(mix[:12345], 512, 384)) < 5e-6


# ------------------------------------------------------------------------------- N1/N2 nets
@pytest.mark.parametrize("mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("kind", ["miso1", "miso3"])
def test_net_golden(kind, mode):
    """MISO_1 / MISO_3 inference against the reference-generated fixture, in the benched tensor-core mode and on the fp32
    FMA path (MISO_3's 2-channel last deconv runs on the tcgen05 kernel too: no FMA fallback inside bf16x3)."""
    from misonet_b200 import synth
    g = _g(f"net_ref_{kind}.npz")
    m, cfg, sd = _model(kind, 0 if kind == "miso1" else 1)
    m.conv_mode = mode
    for b, t in ((2, 20), (1, 11)):
        mix = torch.from_numpy(synth.random_spec(7 + b, (b, 6, t, 129))).cuda()
        with torch.no_grad():
            if kind == "miso1":
                y = m(mix)
            else:
                a2 = torch.from_numpy(synth.random_spec(17 + b, (b, 1, t, 129))).cuda()
                a3 = torch.from_numpy(synth.random_spec(27 + b, (b, 1, t, 129))).cuda()
                y = m(mix, a2, a3)
        assert y.dtype == torch.complex64 and tuple(y.shape) == g[f"y_b{b}"].shape
        errs = {}
        for name in ("enc4", "enc6", "dec0", "dec2"):
            tap = m.tap(name, b, t, 129).cpu().numpy().reshape(g[f"{name}_b{b}"].shape)
            errs[name] = rel_err(tap, g[f"{name}_b{b}"])
        enc0 = m.tap("enc0", b, t, 129).cpu().numpy().reshape(b, 24, t, 127)[:, :, ::3, ::9]
        errs["enc0"] = rel_err(enc0, g[f"enc0_sub_b{b}"])
        tcn = m.tap("tcn", b, t, 129).cpu().numpy().reshape(g[f"tcn_b{b}"].shape)
        errs["tcn"] = rel_err(tcn, g[f"tcn_b{b}"])
        errs["y"] = rel_err(y.cpu().numpy(), g[f"y_b{b}"])
        print(kind, mode, b, t, {k: f"{v:.2e}" for k, v in errs.items()})
        assert max(errs.values()) < 2e-4, errs         # required: REQUIRED_TOL
        assert errs["y"] < REQUIRED_TOL


@pytest.mark.parametrize("mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("kind", ["miso1", "miso3"])
def test_net_paper_golden(kind, mode):
    """PAPER layout (8 blocks, 257 bins, TCN width 384 -- the bench's layout) against the fixture generated from the
    patched copy of the reference's model.py (oracle/make_golden.py:golden_net_paper)."""
    from misonet_b200 import synth
    g = _g("net_ref_paper.npz")
    m, cfg, sd = _model(kind, 3 if kind == "miso1" else 4, layout="PAPER")
    m.conv_mode = mode
    b, t = 2, 9
    mix = torch.from_numpy(synth.random_spec(31, (b, 6, t, 257))).cuda()
    with torch.no_grad():
        if kind == "miso1":
            y = m(mix)
        else:
            a2 = torch.from_numpy(synth.random_spec(32, (b, 1, t, 257))).cuda()
            a3 = torch.from_numpy(synth.random_spec(33, (b, 1, t, 257))).cuda()
            y = m(mix, a2, a3)
    errs = {"y": rel_err(y.cpu().numpy(), g[f"{kind}_y"])}
    for name in ("enc4", "enc7", "dec0", "dec2", "tcn"):
        want = g[f"{kind}_{name}"]
        errs[name] = rel_err(m.tap(name, b, t, 257).cpu().numpy().reshape(want.shape), want)
    enc0 = m.tap("enc0", b, t, 257).cpu().numpy().reshape(b, 24, t, 255)[:, :, ::3, ::17]
    errs["enc0"] = rel_err(enc0, g[f"{kind}_enc0_sub"])
    print(kind, mode, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < 2e-4, errs
    assert errs["y"] < REQUIRED_TOL


def test_net_full_size_vs_oracle():
    """REF shape [1,6,501,129] against the CPU oracle (the oracle takes a few seconds)."""
    from misonet_b200 import synth
    from oracle import miso_net_torch as mnt
    m, cfg, sd = _model("miso1", 0)
    mix = synth.random_spec(3, (1, 6, 501, 129))
    ref = mnt.miso1_forward(sd, cfg, torch.from_numpy(mix)).numpy()
    with torch.no_grad():
        y = m(torch.from_numpy(mix).cuda()).cpu().numpy()
    e = rel_err(y, ref)
    print("full-size MISO_1 rel err", e)
    assert e < 2e-4 and e < REQUIRED_TOL


@pytest.mark.parametrize("mode,tol", [("bf16x3", 2e-4), ("bf16", 5e-2)])
def test_net_tensor_core_modes_vs_oracle(mode, tol):
    """tcgen05 conv path.  bf16x3 (hi/lo split, 3 MMAs per product) must meet north_star's 1e-3 with a wide
    margin; plain bf16 operands are a throughput mode whose error is reported, not held to 1e-3
    (bf16 operands give ~1e-2 vs the fp32 reference)."""
    from misonet_b200 import synth
    from oracle import miso_net_torch as mnt
    m, cfg, sd = _model("miso1", 0)
    m.conv_mode = mode
    mix = synth.random_spec(3, (1, 6, 501, 129))
    ref = mnt.miso1_forward(sd, cfg, torch.from_numpy(mix)).numpy()
    with torch.no_grad():
        y = m(torch.from_numpy(mix).cuda()).cpu().numpy()
    e = rel_err(y, ref)
    print(f"full-size MISO_1 {mode} rel err", e)
    assert e < tol
    if mode == "bf16x3":
        assert e < REQUIRED_TOL
    # ragged T / small batch shapes through the same kernels, PAPER layout
    m2, cfg2, sd2 = _model("miso1", 3, layout="PAPER")
    m2.conv_mode = mode
    mix2 = synth.random_spec(1, (2, 6, 13, 257))
    ref2 = mnt.miso1_forward(sd2, cfg2, torch.from_numpy(mix2)).numpy()
    with torch.no_grad():
        y2 = m2(torch.from_numpy(mix2).cuda()).cpu().numpy()
    assert rel_err(y2, ref2) < tol


def test_net_tensor_core_decisions_match_fp32():
    """The alignment decisions of the shifted MISO1 inference must not change with the conv mode."""
    from misonet_b200 import separation, synth
    m, cfg, sd = _model("miso1", 0)
    mix = torch.from_numpy(synth.random_spec(60, (2, 6, 40, 129))).cuda()
    _, p0 = separation.miso1_inference(m, mix, ref_ch=1, return_perm=True)
    m.conv_mode = "bf16x3"
    _, p1 = separation.miso1_inference(m, mix, ref_ch=1, return_perm=True)
    assert torch.equal(p0, p1)


def test_net_graph_replay_and_weight_update():
    """The forward is replayed as a CUDA graph (include/misonet_b200.h, miso_net_set_graph): replays must be
    bit-identical to the eager launches, and a parameter update must show up without re-capturing."""
    from misonet_b200 import synth
    from oracle import miso_net_torch as mnt
    m, cfg, sd = _model("miso1", 0)
    m.conv_mode = "bf16x3"
    mix = torch.from_numpy(synth.random_spec(21, (2, 6, 24, 129))).cuda()
    with torch.no_grad():
        m.use_graph = False
        y_eager = m(mix).clone()
        m.use_graph = True
        y_cap = m(mix).clone()          # captures
        y_rep = m(mix).clone()          # replays
        # same kernels, same arguments, order-independent (integer) statistics atomics: bitwise reproducible
        assert torch.equal(y_eager, y_cap) and torch.equal(y_eager, y_rep)
        # in-place weight change: same graph, new weights
        key = "encoders.0.0.conv2d.weight"
        sd2 = {k: v.clone() for k, v in sd.items()}
        sd2[key] = sd2[key] * 1.5
        m.load_state_dict(sd2)
        y_new = m(mix)
        ref = mnt.miso1_forward(sd2, cfg, mix.cpu()).numpy()
        assert rel_err(y_new.cpu().numpy(), ref) < 2e-4
        assert rel_err(y_new.cpu().numpy(), y_rep.cpu().numpy()) > 1e-3


def test_net_paper_layout_vs_oracle():
    """8-block / 257-bin / 384-wide-TCN layout (model.py:13-14,30 comments) against the oracle."""
    from misonet_b200 import synth
    from oracle import miso_net_torch as mnt
    m, cfg, sd = _model("miso1", 3, layout="PAPER")
    mix = synth.random_spec(1, (2, 6, 12, 257))
    ref = mnt.miso1_forward(sd, cfg, torch.from_numpy(mix)).numpy()
    with torch.no_grad():
        y = m(torch.from_numpy(mix).cuda()).cpu().numpy()
    assert y.shape == (2, 2, 12, 257)
    assert rel_err(y, ref) < 2e-4


@pytest.mark.parametrize("B,T", [(2, 40), (3, 100), (1, 36)])
def test_net_row_streaming_convs_vs_oracle(B, T):
    """conv_rs.cu (frame taps merged into N): PAPER layout in the tensor-core mode so that the wide stages (255 and 127
    bins: one or two column regions), the packed stages (63 and 31 bins: 2 / 4 frame strips per M tile, T divisible
    by 4) and CTA ranges that start / end inside a strip or cross samples are all exercised."""
    from misonet_b200 import synth
    from oracle import miso_net_torch as mnt
    m, cfg, sd = _model("miso1", 5, layout="PAPER")
    m.conv_mode = "bf16x3"
    mix = synth.random_spec(7, (B, 6, T, 257))
    ref = mnt.miso1_forward(sd, cfg, torch.from_numpy(mix)).numpy()
    with torch.no_grad():
        y = m(torch.from_numpy(mix).cuda()).cpu().numpy()
    e = rel_err(y, ref)
    print(f"PAPER layout B={B} T={T} bf16x3 rel err", e)
    assert e < 2e-4 and e < REQUIRED_TOL


def test_net_bench_workload_vs_oracle():
    """BASELINE configs[1] itself (B = 16 x 6 mics x 500 frames x 257 bins, PAPER layout, bf16x3): three of the sixteen
    samples against the oracle (samples are independent, so the oracle only runs those)."""
    from misonet_b200 import synth
    from oracle import miso_net_torch as mnt
    m, cfg, sd = _model("miso1", 0, layout="PAPER")
    m.conv_mode = "bf16x3"
    mix = synth.random_spec(41, (16, 6, 500, 257))
    with torch.no_grad():
        y = m(torch.from_numpy(mix).cuda()).cpu().numpy()
    assert y.shape == (16, 2, 500, 257)
    for b in (0, 7, 15):
        ref = mnt.miso1_forward(sd, cfg, torch.from_numpy(mix[b:b + 1])).numpy()
        e = rel_err(y[b:b + 1], ref)
        print(f"bench workload sample {b}: rel err {e:.2e}")
        assert e < 2e-4 and e < REQUIRED_TOL


def test_net_batch_invariance_and_chunking():
    """A sample's output must not depend on what else is in the batch (no cross-sample coupling:
    InstanceNorm/gLN only), nor on the workspace-driven batch chunking."""
    from misonet_b200 import synth
    m, cfg, sd = _model("miso1", 0)
    mix = torch.from_numpy(synth.random_spec(11, (3, 6, 16, 129))).cuda()
    with torch.no_grad():
        y_all = m(mix)
        y_one = torch.cat([m(mix[i:i + 1]) for i in range(3)])
        m.max_workspace_bytes = 1            # forces chunks of one sample
        y_chunk = m(mix)
    assert rel_err(y_one.cpu().numpy(), y_all.cpu().numpy()) < 1e-6
    assert rel_err(y_chunk.cpu().numpy(), y_all.cpu().numpy()) < 1e-6
    # the same on the tensor-core path (the tiling, hence the fp32 grouping of the per-CTA statistics partial sums,
    # may depend on the batch size: equal to rounding, not necessarily bitwise)
    m.max_workspace_bytes = 48 << 30
    m.conv_mode = "bf16x3"
    with torch.no_grad():
        z_all = m(mix)
        z_one = torch.cat([m(mix[i:i + 1]) for i in range(3)])
    assert rel_err(z_one.cpu().numpy(), z_all.cpu().numpy()) < 2e-5


def test_net_bad_shape_is_a_clear_error():
    from misonet_b200 import _lib, synth
    m, _, _ = _model("miso1", 0)
    mix = torch.from_numpy(synth.random_spec(1, (1, 6, 8, 257))).cuda()
    with pytest.raises(_lib.MisoError, match="needs F=129"):
        with torch.no_grad():
            m(mix)


# ------------------------------------------------------------------------------- A1 / L1 / L2
def test_losses_golden_and_decisions():
    from misonet_b200 import criterion
    from oracle import miso_np
    g = _g("loss_ref.npz")
    for i in range(2):
        est = torch.from_numpy(g[f"est{i}"]).cuda()
        ref = torch.from_numpy(g[f"ref{i}"]).cuda()
        loss, idx = criterion.loss_uPIT(2, est, [ref[:, 0], ref[:, 1]], return_perm=True)
        assert abs(loss.item() - float(g[f"upit{i}"])) <= 2e-6 * abs(float(g[f"upit{i}"]))
        o_loss, o_idx, o_pair = miso_np.loss_upit(g[f"est{i}"], g[f"ref{i}"])
        assert np.array_equal(idx.cpu().numpy(), o_idx)                 # bit-exact decisions
        assert idx[0].item() == 1
        le = criterion.loss_Enhance(torch.from_numpy(g[f"e1_{i}"]).cuda(), torch.from_numpy(g[f"r1_{i}"]).cuda())
        assert abs(le.item() - float(g[f"enh{i}"])) <= 2e-6 * abs(float(g[f"enh{i}"]))


@pytest.mark.parametrize("S", [2, 3])
def test_pair_decisions_vs_oracle(S):
    from misonet_b200 import criterion, synth
    from oracle import miso_np
    B, T, F = 5, 37, 129
    a = synth.random_spec(200 + S, (B, S, T, F))
    rng = np.random.default_rng(S)
    b = np.stack([a[i, rng.permutation(S)] for i in range(B)]) + 0.1 * synth.random_spec(300 + S, (B, S, T, F))
    pair, idx, _ = criterion.pair_decide(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 0)
    d = miso_np.align_distance(a, b)
    o_idx, _ = miso_np.best_perm(d, S)
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert rel_err(pair.cpu().numpy(), d) < 1e-5
    out = criterion.perm_gather(torch.from_numpy(b).cuda(), idx).cpu().numpy()
    table = miso_np.perm_table(S)
    for i in range(B):
        for s in range(S):
            assert np.array_equal(out[i, s], b[i, table[o_idx[i]][s]])


def test_miso1_inference_golden():
    from misonet_b200 import separation
    g = _g("miso1_inference_ref.npz")
    m, _, _ = _model("miso1", 0)
    for i in range(2):
        out = separation.miso1_inference(m, torch.from_numpy(g[f"mix{i}"]).cuda(), int(g[f"ref_ch{i}"]))
        for k in range(2):
            assert rel_err(out[k].cpu().numpy(), g[f"spk{k}_{i}"]) < 2e-4


def test_miso1_inference_batched_vs_oracle():
    from misonet_b200 import separation, synth
    from oracle import miso_net_torch as mnt
    m, cfg, sd = _model("miso1", 0)
    mix = synth.random_spec(60, (2, 6, 10, 129))
    ref_out, ref_perm = mnt.miso1_inference(sd, cfg, torch.from_numpy(mix), ref_ch=1)
    out, perm = separation.miso1_inference(m, torch.from_numpy(mix).cuda(), ref_ch=1, return_perm=True)
    assert np.array_equal(perm.cpu().numpy(), ref_perm)                # bit-exact alignment decisions
    for k in range(2):
        assert rel_err(out[k].cpu().numpy(), ref_out[k].numpy()) < 2e-4


def test_align_to_clean_vs_oracle():
    from misonet_b200 import separation, synth
    from oracle import miso_np
    B, S, M, T, F = 4, 2, 6, 12, 129
    est = synth.random_spec(70, (S, B, M, T, F))
    clean = est[:, :, 0].transpose(1, 0, 2, 3).copy()
    clean[1] = clean[1, ::-1]
    clean[3] = clean[3, ::-1]
    clean = clean + 0.05 * synth.random_spec(71, clean.shape)
    out, idx = separation.align_to_clean(torch.from_numpy(clean).cuda(), torch.from_numpy(est).cuda(), 0, return_perm=True)
    o_idx, gather = miso_np.clean_align(clean, est[:, :, 0].transpose(1, 0, 2, 3))
    assert np.array_equal(idx.cpu().numpy(), o_idx) and list(o_idx) == [0, 1, 0, 1]
    out = out.cpu().numpy()
    for b in range(B):
        for s in range(S):
            assert np.array_equal(out[s, b], est[gather[b, s], b])


# ------------------------------------------------------------------------------- M1..M7 MVDR
def test_mvdr_golden():
    from misonet_b200 import beamforming
    g = _g("mvdr_ref.npz")
    for i in range(2):
        y = beamforming.Apply_Beamforming(g[f"src{i}"], g[f"mix{i}"])
        assert y.dtype == torch.complex64 and not y.is_cuda and tuple(y.shape) == g[f"y{i}"].shape
        e = rel_err(y.numpy(), g[f"y{i}"])
        print("mvdr golden", i, e)
        assert e < 2e-4 and e < REQUIRED_TOL           # reference itself is complex64: 3e-5..1e-4 noise


def test_mvdr_full_size_vs_oracle_and_distortionless():
    from misonet_b200 import beamforming, synth
    from oracle import miso_np
    B, F, M, T, S = 2, 129, 6, 501, 2
    srcs, mix = [], None
    for s in range(S):
        a, mx = synth.mvdr_case(400 + s, B, F, M, T)
        srcs.append(a)
        mix = mx if mix is None else mix + a
    src_t = torch.stack([torch.from_numpy(a).permute(0, 2, 3, 1) for a in srcs]).contiguous().cuda()   # [S,B,M,T,F]
    mix_t = torch.from_numpy(mix).permute(0, 2, 3, 1).contiguous().cuda()
    y, w = beamforming.mvdr(src_t, mix_t, return_weights=True)
    for s in range(S):
        ref, parts = miso_np.apply_beamforming(srcs[s], mix, return_parts=True)
        assert rel_err(y[s].cpu().numpy(), ref) < 3e-4
        # size-independent property: distortionless response w^H d = 1 for the steering vector
        d = torch.from_numpy(parts["steering"]).cuda()
        resp = (w[s].conj() * d).sum(-1)
        assert torch.allclose(resp, torch.ones_like(resp), atol=2e-3)
    # ragged F (not a multiple of the 32-wide tile) and a different mic count
    a, mx = synth.mvdr_case(410, 1, 45, 4, 77)
    y4 = beamforming.Apply_Beamforming(a, mx).numpy()
    assert rel_err(y4, miso_np.apply_beamforming(a, mx)) < 3e-4


def test_mvdr_paper_size_vs_oracle():
    """MVDR at the bench's PAPER shape: 257 bins x 500 frames, 6 mics, two sources (tester.py:1071-1136)."""
    from misonet_b200 import beamforming, synth
    from oracle import miso_np
    B, F, M, T, S = 1, 257, 6, 500, 2
    srcs, mix = [], None
    for s in range(S):
        a, mx = synth.mvdr_case(500 + s, B, F, M, T)
        srcs.append(a)
        mix = mx if mix is None else mix + a
    src_t = torch.stack([torch.from_numpy(a).permute(0, 2, 3, 1) for a in srcs]).contiguous().cuda()   # [S,B,M,T,F]
    mix_t = torch.from_numpy(mix).permute(0, 2, 3, 1).contiguous().cuda()
    y, w = beamforming.mvdr(src_t, mix_t, return_weights=True)
    for s in range(S):
        ref, parts = miso_np.apply_beamforming(srcs[s], mix, return_parts=True)
        e = rel_err(y[s].cpu().numpy(), ref)
        print("mvdr 257 x 500 source", s, e)
        assert e < 3e-4 and e < REQUIRED_TOL
        d = torch.from_numpy(parts["steering"]).cuda()
        resp = (w[s].conj() * d).sum(-1)
        assert torch.allclose(resp, torch.ones_like(resp), atol=2e-3)      # distortionless response


# ------------------------------------------------------------------------------- pipeline
def test_mvdr_utterance_level_staged():
    """Staged MVDR (tester.py:425-449: covariances over all frames of a chunked recording): one rank == the fused call
    bit for bit; two simulated ranks (frames split in halves, partial sums concatenated along the split axis) agree
    with the fused call to rounding and with the oracle."""
    from misonet_b200 import _lib, beamforming, synth
    from oracle import miso_np
    src_np, mix_np = synth.mvdr_case(5, 1, 129, 6, 400)                    # [B,F,M,T]
    src = torch.from_numpy(src_np).permute(0, 2, 3, 1).contiguous().cuda()    # [B,M,T,F]
    mix = torch.from_numpy(mix_np).permute(0, 2, 3, 1).contiguous().cuda()
    full, w_full = beamforming.mvdr(src.unsqueeze(0), mix, return_weights=True)
    one, w_one = beamforming.mvdr_utterance(src.unsqueeze(0), mix)
    assert torch.equal(one, full) and torch.equal(w_one, w_full)
    lib = _lib.load()
    S, B, M, T, F = 1, 1, 6, 400, 129
    tsplit = lib.miso_mvdr_tsplit(B, F)
    parts, halves = [], [(0, 170), (170, 400)]
    for lo, hi in halves:
        s_h, m_h = src[:, :, lo:hi].contiguous(), mix[:, :, lo:hi].contiguous()
        p = torch.empty(S * B, tsplit, 84, F, dtype=torch.float32, device="cuda")
        _lib.check(lib.miso_mvdr_scm(_lib.ptr(s_h), 0, _lib.ptr(m_h), *m_h.stride(), _lib.ptr(p), S, B, M, hi - lo, F,
                                     _lib.stream_ptr()), "miso_mvdr_scm")
        parts.append(p)
    allp = torch.stack(parts, dim=1).reshape(S * B, 2 * tsplit, 84, F).contiguous()
    w = torch.empty(S, B, F, M, dtype=torch.complex64, device="cuda")
    ws = torch.empty(S * B * F * M * 16, dtype=torch.uint8, device="cuda")
    _lib.check(lib.miso_mvdr_weights(_lib.ptr(allp), 2 * tsplit, T, _lib.ptr(w), S, B, M, F, 1e-6, _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr()), "miso_mvdr_weights")
    outs = []
    for lo, hi in halves:
        m_h = mix[:, :, lo:hi].contiguous()
        o = torch.empty(S, B, hi - lo, F, dtype=torch.complex64, device="cuda")
        _lib.check(lib.miso_mvdr_apply(_lib.ptr(m_h), *m_h.stride(), _lib.ptr(w), _lib.ptr(o), S, B, M, hi - lo, F,
                                       _lib.stream_ptr()), "miso_mvdr_apply")
        outs.append(o)
    two = torch.cat(outs, dim=2)
    assert rel_err(two.cpu().numpy(), full.cpu().numpy()) < 1e-5
    ref = miso_np.apply_beamforming(src_np, mix_np)                            # [B,T,F]
    assert rel_err(two[0].cpu().numpy(), ref) < REQUIRED_TOL


def test_pipeline_full_size_vs_oracle():
    """One synthetic SMS-WSJ-shaped 4 s utterance through STFT -> MISO1 x6 -> align -> MVDR x2 ->
    MISO3 x2 against the oracle pipeline fed stage by stage (MVDR amplifies
    upstream differences when the eigen-gap is small, so stages are checked with oracle-fed inputs)."""
    from misonet_b200 import synth, pipeline
    from oracle import miso_np
    from oracle import miso_net_torch as mnt
    m1, cfg1, sd1 = _model("miso1", 0)
    m3, cfg3, sd3 = _model("miso3", 1)
    mix_t, _ = synth.make_utterance(5, n_samples=8000)               # 1 s keeps the CPU oracle fast
    pipe = pipeline.MisoBfMiso(m1, m3)
    out = pipe(torch.from_numpy(mix_t[None]).cuda())
    mix_stft = miso_np.stft(mix_t)[None]                             # [1,6,T,129]
    assert rel_err(out["mix_stft"].cpu().numpy(), mix_stft) < 5e-6
    o_miso1, _ = mnt.miso1_inference(sd1, cfg1, torch.from_numpy(mix_stft), 0)
    got_miso1 = out["miso1"].cpu().numpy()
    for s in range(2):
        assert rel_err(got_miso1[s], o_miso1[s].numpy()) < 2e-4
    # MVDR on OUR miso1 output, oracle-fed
    for s in range(2):
        src = np.transpose(got_miso1[s], (0, 3, 1, 2))
        ref_bf = miso_np.apply_beamforming(src, np.transpose(mix_stft, (0, 3, 1, 2)))
        assert rel_err(out["beamformed"][s].cpu().numpy(), ref_bf) < 1e-3
        ref_enh = mnt.miso3_forward(sd3, cfg3, torch.from_numpy(mix_stft), out["beamformed"][s].cpu().unsqueeze(1),
                                    torch.from_numpy(got_miso1[s][:, 0:1])).numpy()
        assert rel_err(out["enhanced"][:, s].cpu().numpy(), ref_enh[:, 0]) < 2e-4


def test_pipeline_paper_shape_bf16x3_vs_oracle():
    """BASELINE configs[2] at the PAPER shape, one utterance, in the benched mode: 512-point STFT (257 bins) x 500 frames
    -> MISO1 x 6 shifts -> MVDR x 2 -> MISO3 x 2, stage by stage against the oracle fed with our upstream outputs."""
    from misonet_b200 import synth, pipeline
    from oracle import miso_np
    from oracle import miso_net_torch as mnt
    m1, cfg1, sd1 = _model("miso1", 0, layout="PAPER")
    m3, cfg3, sd3 = _model("miso3", 1, layout="PAPER")
    assert m1.conv_mode == "bf16x3" and m3.conv_mode == "bf16x3"      # the default IS the benched mode
    mix_t, _ = synth.make_utterance(9, n_samples=499 * 128)
    pipe = pipeline.MisoBfMiso(m1, m3, nperseg=512, noverlap=384)
    out = pipe(torch.from_numpy(mix_t[None]).cuda())
    mix_stft = miso_np.stft(mix_t, 512, 384)[None]
    assert mix_stft.shape == (1, 6, 500, 257)
    assert rel_err(out["mix_stft"].cpu().numpy(), mix_stft) < 5e-6
    o_miso1, o_perm = mnt.miso1_inference(sd1, cfg1, torch.from_numpy(mix_stft), 0)
    got_miso1 = out["miso1"].cpu().numpy()
    for s in range(2):
        assert rel_err(got_miso1[s], o_miso1[s].numpy()) < 2e-4
    for s in range(2):
        src = np.transpose(got_miso1[s], (0, 3, 1, 2))
        ref_bf = miso_np.apply_beamforming(src, np.transpose(mix_stft, (0, 3, 1, 2)))
        assert rel_err(out["beamformed"][s].cpu().numpy(), ref_bf) < 1e-3
        ref_enh = mnt.miso3_forward(sd3, cfg3, torch.from_numpy(mix_stft), out["beamformed"][s].cpu().unsqueeze(1),
                                    torch.from_numpy(got_miso1[s][:, 0:1])).numpy()
        e = rel_err(out["enhanced"][:, s].cpu().numpy(), ref_enh[:, 0])
        print("pipeline PAPER bf16x3 enhanced", s, e)
        assert e < 2e-4


def test_module_on_a_non_current_device():
    """model.cuda(1) with cuda:0 current (the reference does model.cuda(gpu_num) without set_device, run.py:68): every
    launch, the graph capture and the workspace must go to the module's device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from misonet_b200 import synth
    from oracle import miso_net_torch as mnt
    m, cfg, sd = _model("miso1", 0)
    m = m.cuda(1)
    torch.cuda.set_device(0)
    mix = synth.random_spec(21, (2, 6, 24, 129))
    ref = mnt.miso1_forward(sd, cfg, torch.from_numpy(mix)).numpy()
    with torch.no_grad():
        for _ in range(2):                                       # capture, then replay
            y = m(torch.from_numpy(mix).to("cuda:1"))
    assert y.device == torch.device("cuda", 1) and torch.cuda.current_device() == 0
    assert rel_err(y.cpu().numpy(), ref) < 2e-4


def test_chunked_batches_keep_the_input_alignment():
    """REF shape (501 x 129: an odd T * F makes every odd sample of the input planes 64-byte aligned only): a workspace
    cap that forces chunking must still give the unchunked result."""
    from misonet_b200 import synth
    m, cfg, sd = _model("miso1", 0)
    mix = torch.from_numpy(synth.random_spec(12, (5, 6, 21, 129))).cuda()
    with torch.no_grad():
        y_all = m(mix).clone()
        per = m._ws_bytes(1, 21, 129)
        m.max_workspace_bytes = 3 * per + per // 2            # would chunk by 3 -> must round to 2
        assert m._chunk(5, 21, 129) == 2
        y_chunk = m(mix)
    assert rel_err(y_chunk.cpu().numpy(), y_all.cpu().numpy()) < 2e-5


def test_long_recording_chunks_and_sharding():
    """continuous.separate_recording (dataloader/data.py:524-597 + tester.py:857-974): equals the per-chunk pipeline
    followed by ISTFT and the gap trim, and the block-partitioned two-rank run is bit-identical to the one-rank run."""
    from misonet_b200 import audio, continuous, synth
    from misonet_b200.pipeline import MisoBfMiso
    m1, _, _ = _model("miso1", 0)
    m3, _, _ = _model("miso3", 1)
    # this test is about the chunk / shard bookkeeping and compares runs with DIFFERENT batch compositions bit for bit:
    # the fp32 path is batch-invariant bitwise, the tensor-core path only to rounding (test_net_batch_invariance_and_chunking)
    m1.conv_mode = m3.conv_mode = "fp32"
    pipe = MisoBfMiso(m1, m3)
    chunk = 64 * 24                                        # 25 frames per chunk
    n = 3 * chunk - 500
    wav = torch.from_numpy(np.ascontiguousarray(synth.make_utterance(21, n_samples=n)[0], dtype=np.float32)).cuda()
    full = continuous.separate_recording(pipe, wav, chunk)
    assert full.shape == (2, n)
    chunks, gap = continuous.chunk_signal(wav, chunk)
    assert gap == 500
    ref = torch.cat([audio.istft(pipe(chunks[i:i + 1])["enhanced"])[0] for i in range(3)], dim=-1)[:, :n]
    assert rel_err(full.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    parts = [continuous.separate_recording(pipe, wav, chunk, rank=r, world=2, gather=False) for r in range(2)]
    both = torch.cat(parts, dim=0).permute(1, 0, 2).reshape(2, -1)[:, :n]
    assert torch.equal(both, full)
    as16 = continuous.separate_recording(pipe, wav, chunk, to_int16=True)
    assert as16.dtype == torch.int16 and as16.shape == (2, n)


def test_utterance_wise_beamforming_of_a_recording():
    """continuous.beamform_recording (tester.py:340-449 with utterance_flag): equals the same composition done by hand
    with the oracle's STFT / MVDR on the MISO1 images."""
    from misonet_b200 import continuous, synth
    from oracle import miso_np
    m1, _, _ = _model("miso1", 0)
    chunk = 64 * 24
    n = 2 * chunk + 700
    mix, images = synth.make_utterance(31, n_samples=n)
    wav = torch.from_numpy(mix).cuda()
    clean = torch.from_numpy(np.ascontiguousarray(images[:, :, 0])).cuda()          # [Spk, N] at the reference mic
    res = continuous.beamform_recording(m1, wav, chunk, clean=clean)
    t_all = miso_np.stft_num_frames(n)
    assert res["beamformed"].shape == (2, t_all, 129) and res["miso1_wav"].shape == (2, 6, n)
    img = res["miso1_wav"].cpu().numpy()                                             # [Spk, Mic, N]
    mix_stft = miso_np.stft(mix)                                                     # [Mic, T, F]
    for s in range(2):
        src_stft = miso_np.stft(img[s].T)                                            # [Mic, T, F]
        ref = miso_np.apply_beamforming(src_stft.transpose(2, 0, 1)[None], mix_stft.transpose(2, 0, 1)[None])[0]   # [T, F]
        assert rel_err(res["beamformed"][s].cpu().numpy(), ref) < REQUIRED_TOL
    assert res["wav"].shape == (2, (t_all - 1) * 64)


def test_dropin_methods():
    """The reference-named methods of B200HotPath at B = 1."""
    from misonet_b200 import dropin
    g = _g("miso1_inference_ref.npz")
    gm = _g("mvdr_ref.npz")
    m1, _, _ = _model("miso1", 0)

    class T(dropin.B200HotPath):
        pass

    t = T()
    t.model_sep, t.num_spks, t.ref_ch, t.device = m1, 2, 0, 0
    out = t.MISO1_Inference(torch.from_numpy(g["mix0"]), ref_ch=0)
    assert not out[0].is_cuda and rel_err(out[0].numpy(), g["spk0_0"]) < 2e-4
    y = t.Apply_Beamforming(gm["src0"], gm["mix0"])
    assert rel_err(y.numpy(), gm["y0"]) < 2e-4


def test_native_library_is_what_ran():
    from misonet_b200 import _lib
    assert _lib.launch_count() > 0
    with open("/proc/self/maps") as f:
        assert "libmisonet_b200.so" in f.read()


def test_fused_tcn_switch():
    """MISO_TCN_FUSED=1 (tcn.cu, tcn_fused_kernel: the whole TCN as one launch, a thread-block cluster per sample; opt-in,
    measured no faster -- profiles/r2_tcn_fused_*) must give the default path's result at one and at several frame tiles per
    sample, both layouts.  The switch is read once per process, hence the subprocess."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from test_gpu_parity import _model
from conftest import rel_err
from misonet_b200 import synth, _lib
from oracle import miso_net_torch as mnt
for layout, F, B, T in (("REF", 129, 1, 200), ("PAPER", 257, 2, 300), ("PAPER", 257, 1, 40)):
    m, cfg, sd = _model("miso1", 5, layout=layout)
    mix = synth.random_spec(7, (B, 6, T, F))
    ref = mnt.miso1_forward(sd, cfg, torch.from_numpy(mix)).numpy()
    with torch.no_grad():
        y = m(torch.from_numpy(mix).cuda()).cpu().numpy()
    e = rel_err(y, ref)
    print("FUSED_TCN_REL_ERR", layout, T, e)
    assert e < 2e-4
''' % (root, root)
    env = dict(os.environ, MISO_TCN_FUSED="1", MISO_TC_DEBUG="1")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("FUSED_TCN_REL_ERR") == 3
