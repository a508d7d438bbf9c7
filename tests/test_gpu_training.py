"""GPU parity tests of the training path (reference trainer.py:159-212):
``estimate = model(mix)``, ``loss = loss_uPIT(...)``, ``loss.backward()`` through the C ABI
(``miso_net_forward_train`` / ``miso_net_backward`` / ``miso_upit_bwd``) against torch autograd over the CPU oracle
(``oracle/miso_net_torch.net_forward``, pinned to the reference through tests/golden) with identical weights and inputs.

Tolerance: north_star's 1e-3 relative on values, applied per parameter tensor to the gradients for a SMOOTH upstream
gradient, against the oracle evaluated in float64.  One caveat is inherent to the network, not to the implementation:
PReLU (model.py:557, slope 0.25) has a kink, and the CUDA forward agrees with the oracle to 2e-5 (bf16 hi/lo activation
storage), so a fraction ~1e-5 of the TCN's pre-activations sit on the other side of zero; each such element changes one
component of the back-propagated gradient by 75 %, which at these small test sizes is 1e-3 ... 1e-2 of the encoder / TCN
gradients (measured; the reference's own float64 gradients move by the same amount when its input is perturbed by
2.5e-5).  The per-parameter test therefore runs with the PReLU slopes set to 1 (every other operation is C1), where all
268 parameter gradients agree to ~1e-4; with the reference's slope 0.25 the decoder gradients (which do not pass through
the TCN) are still held to 2e-4 and the rest to an aggregate bound.  The L1 losses of criterion.py have a sign() in
their gradient for the same reason: that test checks the loss-gradient kernel on identical inputs (tight) and the
end-to-end parameter gradients by cosine similarity.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

REQUIRED_TOL = 1e-3


def _model(seed, layout="REF", mode="bf16x3", prelu_alpha=None):
    from misonet_b200.model import MISO_1
    from oracle import weights
    from oracle import miso_net_torch as mnt
    en, de = mnt.LAYOUTS[layout]
    cfg = mnt.NetConfig.miso1(layout=layout)
    m = MISO_1(2, 6, len(en), list(en), list(de), "IN")
    sd = weights.make_state_dict(cfg, seed)
    if prelu_alpha is not None:
        for k in sd:
            if k.startswith("TCN.") and k.endswith(".net.1.weight"):
                sd[k] = torch.full_like(sd[k], prelu_alpha)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    m.conv_mode = mode
    return m, cfg, sd


def _oracle_grads(sd, cfg, mix, upstream=None, refs=None, dtype=torch.float32, perturb=0.0):
    """torch autograd over the oracle network; returns (output complex [B,S,T,F], {key: grad}, loss or None)."""
    from oracle import miso_net_torch as mnt
    sdr = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    x = (torch.cat((mix.real, mix.imag), dim=1) if mix.is_complex() else mix).to(dtype)   # a real tensor is the network input itself
    if perturb:
        x = x * (1.0 + perturb * torch.randn(x.shape, dtype=dtype, generator=torch.Generator().manual_seed(99)))
    if upstream is not None:
        upstream = upstream.to(torch.complex128 if dtype == torch.float64 else torch.complex64)
    y = mnt.net_forward(sdr, cfg, x)
    S = y.shape[1] // 2
    loss = None
    if upstream is not None:
        L = (y[:, :S] * upstream.real + y[:, S:] * upstream.imag).sum()
    else:
        import itertools
        est = torch.complex(y[:, :S], y[:, S:])
        e, r = est.unsqueeze(2), refs.unsqueeze(1)
        pair = ((e.real - r.real).abs().sum((3, 4)) + (e.imag - r.imag).abs().sum((3, 4)) +
                (torch.sqrt(e.real ** 2 + e.imag ** 2 + 1e-8) - r.abs()).abs().sum((3, 4)))          # criterion.py:27-33
        perms = list(itertools.permutations(range(S)))
        per = torch.stack([sum(pair[:, i, p[i]] for i in range(S)) for p in perms], dim=1)
        L = per.min(dim=1).values.mean()
        loss = L.detach()
    L.backward()
    out = torch.complex(y[:, :S], y[:, S:]).detach()
    return out, {k: v.grad for k, v in sdr.items()}, loss


FORWARD_AGREEMENT = 2.5e-5     # measured forward error of the CUDA path against the oracle (bf16 hi/lo storage)


def _check_grads(m, ref_grads, what, tol=REQUIRED_TOL, kinked=False):
    """Per-parameter relative error against float64 autograd.  kinked = True (PReLU slope 0.25): only the decoders are
    held to the per-parameter bound, everything to an aggregate one (module docstring)."""
    scale = max(float(v.norm()) for v in ref_grads.values())
    num = den = 0.0
    failures = []
    scal_g, scal_r = [], []     # the 28 PReLU slopes are scalars with heavy cancellation: compared as one vector
    for k, p in m.named_parameters():
        assert p.grad is not None, f"{what}: no gradient for {k}"
        g, r = p.grad.detach().cpu().double().numpy().ravel(), ref_grads[k].numpy().ravel()
        assert np.isfinite(g).all(), f"{what}: non-finite gradient for {k}"
        num += float(((g - r) ** 2).sum())
        den += float((r ** 2).sum())
        if r.size == 1:
            scal_g.append(g[0])
            scal_r.append(r[0])
            continue
        if float(np.linalg.norm(r)) < 1e-7 * scale:
            # mathematically zero gradient (gLN beta ahead of an InstanceNorm1d): rounding noise on both sides
            assert float(np.linalg.norm(g)) < 1e-6 * scale, f"{what}: {k} should have a vanishing gradient"
            continue
        e = rel_err(g, r)
        bound = 2e-4 if k.startswith("decoders.") else (3e-2 if kinked else tol)
        if e > bound:
            failures.append((k, e))
    e = rel_err(np.array(scal_g), np.array(scal_r))
    if e > (1e-1 if kinked else tol):
        failures.append(("PReLU slopes (pooled)", e))
    total = (num / den) ** 0.5
    assert not failures, f"{what}: (parameter, rel err) {failures[:8]} (all parameters {total:.3e})"
    assert total < (1e-2 if kinked else 2e-4), f"{what}: all-parameter error {total:.3e}"
    return total


@pytest.mark.parametrize("mode,B,T,seed", [("bf16x3", 2, 40, 3), ("fp32", 1, 24, 4)])
def test_backward_matches_autograd_ref_layout(mode, B, T, seed):
    """Every parameter gradient of MISO_1 (REF layout, F=129) for a smooth upstream gradient, PReLU slopes 1."""
    from misonet_b200 import synth
    m, cfg, sd = _model(seed, "REF", mode, prelu_alpha=1.0)
    mix = torch.from_numpy(synth.random_spec(11, (B, 6, T, 129)))
    up = torch.from_numpy(synth.random_spec(12, (B, 2, T, 129)))
    out_ref, g_ref, _ = _oracle_grads(sd, cfg, mix, upstream=up, dtype=torch.float64)
    out = m(mix.cuda())
    assert out.requires_grad
    assert rel_err(out.detach().cpu().numpy(), out_ref.numpy()) < 2 * FORWARD_AGREEMENT
    out.backward(up.cuda())
    _check_grads(m, g_ref, f"REF {mode}")
    # a second step reuses the workspace and must give the same gradients (atomics: equal to rounding, not bitwise)
    first = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    m(mix.cuda()).backward(up.cuda())
    scale = max(float(g.norm()) for g in first.values())
    for k, p in m.named_parameters():
        # (a PReLU slope's gradient is ONE scalar summed over every TCN activation, with heavy cancellation: looser)
        rel = 5e-3 if p.numel() == 1 else 1e-4   # (observed up to 6e-4 on a slope gradient; its accuracy is held separately above)
        assert float((p.grad - first[k]).norm()) < rel * float(first[k].norm()) + 1e-8 * scale, k


def test_backward_matches_autograd_paper_layout():
    """PAPER layout (8 blocks, F=257, TCN width 384) at a short T, PReLU slopes 1."""
    from misonet_b200 import synth
    m, cfg, sd = _model(5, "PAPER", "bf16x3", prelu_alpha=1.0)
    B, T = 1, 20
    mix = torch.from_numpy(synth.random_spec(21, (B, 6, T, 257)))
    up = torch.from_numpy(synth.random_spec(22, (B, 2, T, 257)))
    _, g_ref, _ = _oracle_grads(sd, cfg, mix, upstream=up, dtype=torch.float64)
    m(mix.cuda()).backward(up.cuda())
    _check_grads(m, g_ref, "PAPER bf16x3")


def test_backward_reference_prelu_slope():
    """The reference's initial PReLU slope 0.25 (kinked): decoders tight, the rest to the aggregate bound."""
    from misonet_b200 import synth
    m, cfg, sd = _model(4, "REF", "bf16x3")
    B, T = 2, 40
    mix = torch.from_numpy(synth.random_spec(11, (B, 6, T, 129)))
    up = torch.from_numpy(synth.random_spec(12, (B, 2, T, 129)))
    _, g_ref, _ = _oracle_grads(sd, cfg, mix, upstream=up, dtype=torch.float64)
    m(mix.cuda()).backward(up.cuda())
    _check_grads(m, g_ref, "REF bf16x3, slope 0.25", kinked=True)
    a = torch.cat([p.grad.flatten() for p in m.parameters()]).cpu().double()
    b = torch.cat([g_ref[k].flatten() for k, _ in m.named_parameters()])
    assert float((a @ b) / (a.norm() * b.norm())) > 0.99995


def test_miso3_backward_and_loss_enhance():
    """MISO_3 training step (trainer.py:398-443): 16 input / 2 output channels, loss_Enhance; PReLU slopes 1 for the
    per-parameter check (module docstring)."""
    from misonet_b200 import criterion, synth
    from misonet_b200.model import MISO_3
    from oracle import weights
    from oracle import miso_net_torch as mnt
    en, de = mnt.LAYOUTS["REF"]
    cfg = mnt.NetConfig.miso3(layout="REF")
    sd = weights.make_state_dict(cfg, 6)
    for k in sd:
        if k.startswith("TCN.") and k.endswith(".net.1.weight"):
            sd[k] = torch.ones_like(sd[k])
    m = MISO_3(1, 6, len(en), list(en), list(de), "IN")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    m.conv_mode = "bf16x3"
    B, T, F = 2, 24, 129
    mix = torch.from_numpy(synth.random_spec(51, (B, 6, T, F)))
    bf = torch.from_numpy(synth.random_spec(52, (B, 1, T, F)))
    m1 = torch.from_numpy(synth.random_spec(53, (B, 1, T, F)))
    up = torch.from_numpy(synth.random_spec(54, (B, 1, T, F)))
    x = torch.cat((mix.real, bf.real, m1.real, mix.imag, bf.imag, m1.imag), dim=1)          # model.py:358-366
    out_ref, g_ref, _ = _oracle_grads(sd, cfg, x, upstream=up, dtype=torch.float64)
    out = m(mix.cuda(), bf.cuda(), m1.cuda())
    assert rel_err(out.detach().cpu().numpy(), out_ref.numpy()) < 2 * FORWARD_AGREEMENT
    out.backward(up.cuda())
    _check_grads(m, g_ref, "MISO_3 REF bf16x3")
    # loss_Enhance and its gradient on identical inputs (criterion.py:121-141)
    est = torch.from_numpy(synth.random_spec(55, (B, 1, T, F)))
    ref = torch.from_numpy(synth.random_spec(56, (B, 1, T, F)))
    e = est.clone().requires_grad_(True)
    L = ((e.real - ref.real).abs().sum() + (e.imag - ref.imag).abs().sum() +
         (torch.sqrt(e.real ** 2 + e.imag ** 2 + 1e-8) - ref.abs()).abs().sum()) / B
    (2.0 * L).backward()
    ec = est.cuda().requires_grad_(True)
    loss = criterion.loss_Enhance(ec, ref.cuda())
    assert abs(loss.item() - L.item()) <= 2e-6 * abs(L.item())
    (2.0 * loss).backward()
    assert rel_err(ec.grad.cpu().numpy(), e.grad.numpy()) < 1e-6
    # one full step: the loss decreases along the negative gradient
    m.zero_grad(set_to_none=True)
    tgt = torch.from_numpy(synth.random_spec(57, (B, 1, T, F))).cuda()
    l0 = criterion.loss_Enhance(m(mix.cuda(), bf.cuda(), m1.cuda()), tgt)
    l0.backward()
    gn2 = sum(float((p.grad ** 2).sum()) for p in m.parameters())
    step = 1e-3 * l0.item() / gn2
    with torch.no_grad():
        for p in m.parameters():
            p -= step * p.grad
        l1 = criterion.loss_Enhance(m(mix.cuda(), bf.cuda(), m1.cuda()), tgt)
    assert l1.item() < l0.item(), (l0.item(), l1.item())


def test_upit_loss_gradient_kernel():
    """miso_upit_bwd against autograd of criterion.py:8-63's arithmetic on the SAME estimate (signs identical)."""
    from misonet_b200 import criterion, synth
    B, S, T, F = 3, 2, 50, 129
    est = torch.from_numpy(synth.random_spec(31, (B, S, T, F)))
    ref = torch.from_numpy(synth.random_spec(32, (B, S, T, F)))
    ref[1] = ref[1].flip(0)                     # make utterance 1 prefer the swapped permutation
    est[1] = ref[1].flip(0) + 0.1 * est[1]
    e = est.clone().requires_grad_(True)
    ee, rr = e.unsqueeze(2), ref.unsqueeze(1)
    pair = ((ee.real - rr.real).abs().sum((3, 4)) + (ee.imag - rr.imag).abs().sum((3, 4)) +
            (torch.sqrt(ee.real ** 2 + ee.imag ** 2 + 1e-8) - rr.abs()).abs().sum((3, 4)))
    per = torch.stack([pair[:, 0, 0] + pair[:, 1, 1], pair[:, 0, 1] + pair[:, 1, 0]], dim=1)
    L = per.min(dim=1).values.mean()
    (3.0 * L).backward()
    ec = est.cuda().requires_grad_(True)
    loss, idx = criterion.loss_uPIT(2, ec, [ref[:, 0].cuda(), ref[:, 1].cuda()], return_perm=True)
    assert np.array_equal(idx.cpu().numpy(), per.argmin(dim=1).numpy())          # bit-exact decision
    assert idx[1].item() == 1
    assert abs(loss.item() - L.item()) <= 2e-6 * abs(L.item())
    (3.0 * loss).backward()
    assert rel_err(ec.grad.cpu().numpy(), e.grad.numpy()) < 1e-6


def test_training_step_upit_end_to_end():
    """trainer.py:159-212 in miniature: forward, loss_uPIT, backward, Adam step; loss and update direction against the oracle."""
    from misonet_b200 import criterion, synth
    m, cfg, sd = _model(7, "REF", "bf16x3")
    B, T = 2, 32
    mix = torch.from_numpy(synth.random_spec(41, (B, 6, T, 129)))
    refs = torch.from_numpy(synth.random_spec(42, (B, 2, T, 129)))
    _, g_ref, loss_ref = _oracle_grads(sd, cfg, mix, refs=refs)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.zero_grad()
    est = m(mix.cuda())
    loss = criterion.loss_uPIT(2, est, [refs[:, 0].cuda(), refs[:, 1].cuda()])
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    loss.backward()
    a = torch.cat([p.grad.flatten() for p in m.parameters()]).cpu().double()
    b = torch.cat([g_ref[k].flatten() for k, _ in m.named_parameters()]).double()
    cos = float((a @ b) / (a.norm() * b.norm()))
    assert cos > 0.9995, f"gradient direction cosine {cos}"
    assert abs(float(a.norm() / b.norm()) - 1.0) < 1e-2
    # an inference forward (validation) between a training forward and its backward leaves the gradients untouched
    g_first = torch.cat([p.grad.flatten() for p in m.parameters()]).clone()
    m.zero_grad(set_to_none=True)
    est = m(mix.cuda())
    with torch.no_grad():
        m(torch.from_numpy(synth.random_spec(43, (B, 6, T, 129))).cuda())
    criterion.loss_uPIT(2, est, [refs[:, 0].cuda(), refs[:, 1].cuda()]).backward()
    g_again = torch.cat([p.grad.flatten() for p in m.parameters()])
    assert float((g_again - g_first).norm() / g_first.norm()) < 1e-4
    torch.nn.utils.clip_grad_norm_(m.parameters(), 10.0)          # trainer.py:210
    opt.step()
    with torch.no_grad():                                         # repacked weights are picked up by the next forward
        est2 = m(mix.cuda())
    loss2 = criterion.loss_uPIT(2, est2, [refs[:, 0].cuda(), refs[:, 1].cuda()])
    assert torch.isfinite(loss2) and loss2.item() != loss.item()


@pytest.mark.parametrize("kind", ["miso1", "miso3"])
def test_training_step_against_reference_fixture_unit_prelu_slopes_loose_bounds(kind):
    """The CUDA training step against the REAL reference's model(mix) -> loss -> loss.backward() (tests/golden/train_ref.npz,
    generated by oracle/make_golden.py:golden_training from model.py + criterion.py; PReLU slopes 1).  The L1 losses'
    sign() makes a single estimate/reference crossing worth ~1e-2 of the gradient at this size (12 k output values), so the
    bounds here are loose by design; the tight gradient checks are the smooth-upstream tests above."""
    import os
    from conftest import GOLDEN
    from misonet_b200 import criterion, synth
    from misonet_b200.model import MISO_1, MISO_3
    from oracle import weights
    from oracle import miso_net_torch as mnt
    g = np.load(os.path.join(GOLDEN, "train_ref.npz"))
    en, de = mnt.LAYOUTS["REF"]
    cfg = mnt.NetConfig.miso1() if kind == "miso1" else mnt.NetConfig.miso3()
    sd = weights.make_state_dict(cfg, 0 if kind == "miso1" else 1)
    for k in sd:
        if k.startswith("TCN.") and k.endswith(".net.1.weight"):
            sd[k] = torch.ones_like(sd[k])
    m = (MISO_1(2, 6, 7, list(en), list(de), "IN") if kind == "miso1" else MISO_3(1, 6, 7, list(en), list(de), "IN"))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    m.conv_mode = "bf16x3"
    b, t = 2, 12
    mix = torch.from_numpy(synth.random_spec(61, (b, 6, t, 129))).cuda()
    if kind == "miso1":
        refs = torch.from_numpy(synth.random_spec(62, (b, 2, t, 129))).cuda()
        est = m(mix)
        loss = criterion.loss_uPIT(2, est, [refs[:, 0], refs[:, 1]])
    else:
        a2 = torch.from_numpy(synth.random_spec(63, (b, 1, t, 129))).cuda()
        a3 = torch.from_numpy(synth.random_spec(64, (b, 1, t, 129))).cuda()
        refs = torch.from_numpy(synth.random_spec(65, (b, 1, t, 129))).cuda()
        est = m(mix, a2, a3)
        loss = criterion.loss_Enhance(est, refs)
    loss.backward()
    assert rel_err(est.detach().cpu().numpy(), g[f"{kind}_est"]) < 2 * FORWARD_AGREEMENT
    assert abs(loss.item() - float(g[f"{kind}_loss"])) <= 1e-4 * abs(float(g[f"{kind}_loss"]))
    ref_norms = g[f"{kind}_grad_norms"]
    norms = np.array([float(p.grad.norm()) for p in m.parameters()])
    numel = np.array([p.numel() for p in m.parameters()])
    big = (ref_norms > 1e-6 * ref_norms.max()) & (numel > 1)       # scalars (PReLU slopes) cancel heavily: covered by the pooled check above
    dev = np.abs(norms - ref_norms) / np.maximum(ref_norms, 1e-30)
    names = [k for k, _ in m.named_parameters()]
    bad = [(names[i], float(dev[i])) for i in np.argsort(-dev * big)[:4] if big[i] and dev[i] > 3e-2]
    assert not bad, bad
    worst = 0.0
    for key in g.files:
        if key.startswith(f"{kind}_grad::"):
            k = key.split("::", 1)[1]
            worst = max(worst, rel_err(dict(m.named_parameters())[k].grad.cpu().numpy(), g[key]))
    assert worst < 3e-2, worst
    print(f"{kind}: worst stored gradient tensor vs the reference {worst:.2e}")


def test_training_full_size_properties():
    """BASELINE configs[3] utterance shape (PAPER layout, 500 frames x 257 bins) where the oracle's autograd is too slow
    to be the checker: size-independent properties of the backward pass.
    (1) The last deconv has no ELU / norm behind it (model.py:418-423), so its bias gradient is exactly the sum of the
        upstream gradient per output channel.
    (2) Directional derivative: for the smooth functional L(theta) = Re <g, model(mix; theta)> the central difference
        along the normalised gradient direction d equals |grad| (grad . d)."""
    from misonet_b200 import synth
    m, cfg, sd = _model(9, "PAPER", "bf16x3", prelu_alpha=1.0)
    B, T, F = 2, 500, 257
    mix = torch.from_numpy(synth.random_spec(71, (B, 6, T, F))).cuda()
    up = torch.from_numpy(synth.random_spec(72, (B, 2, T, F))).cuda()
    out = m(mix)
    out.backward(up)
    params = dict(m.named_parameters())
    gb = params["decoders.7.1.deconv2d.bias"].grad.cpu().double().numpy()
    want = torch.cat((up.real.sum(dim=(0, 2, 3)), up.imag.sum(dim=(0, 2, 3)))).cpu().double().numpy()      # channels re(s0, s1), im(s0, s1)
    assert rel_err(gb, want) < 1e-4
    grads = [p.grad.clone() for p in m.parameters()]
    assert all(torch.isfinite(g).all() for g in grads)
    gnorm = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads)))

    def functional():
        with torch.no_grad():
            o = m(mix)
        return float((o.real.double() * up.real.double() + o.imag.double() * up.imag.double()).sum())

    eps = 2e-3
    with torch.no_grad():
        for p, g in zip(m.parameters(), grads):
            p.add_(g, alpha=eps / gnorm)
        lp = functional()
        for p, g in zip(m.parameters(), grads):
            p.add_(g, alpha=-2 * eps / gnorm)
        lm = functional()
        for p, g in zip(m.parameters(), grads):
            p.add_(g, alpha=eps / gnorm)
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - gnorm) < 2e-2 * gnorm, (fd, gnorm)


def test_training_graph_replay_and_gradient_buckets():
    """Round-2 training plumbing.  (1) From the third step with the same buffers the forward and the backward are replayed as
    CUDA graphs: the gradient of a replayed step equals the eager one to rounding (the tcgen05 weight gradient reduces its
    per-CTA partial sums in a fixed order; the fp32 atomics of the bias / gLN / depthwise gradient kernels and of the
    InstanceNorm backward sums are order dependent: ~1e-5 run to run, measured).
    (2) miso_net_grad_buckets: five contiguous ranges in completion order that tile the flat gradient buffer exactly, and
    miso_net_wait_grad_bucket accepts every one of them after a backward."""
    import ctypes
    from misonet_b200 import _lib, synth
    m, cfg, sd = _model(0, "REF", "bf16x3", 1.0)
    mix = torch.from_numpy(synth.random_spec(3, (2, 6, 24, 129))).cuda()
    up = torch.from_numpy(synth.random_spec(4, (2, 2, 24, 129))).cuda()
    grads = []
    for it in range(4):          # eager, capture, replay, replay
        m.zero_grad(set_to_none=True)
        est = m(mix)
        (est.real * up.real + est.imag * up.imag).sum().backward()
        grads.append(torch.cat([p.grad.flatten() for p in m.parameters()]).clone())
    scale = float(grads[0].norm())
    for g in grads[1:]:
        assert float((g - grads[0]).norm()) <= 1e-4 * scale
    lib = _lib.load()
    b0, b1 = (ctypes.c_int64 * 8)(), (ctypes.c_int64 * 8)()
    nb = lib.miso_net_grad_buckets(m._handle, b0, b1, 8)
    assert nb == 5
    ranges = sorted((int(b0[k]), int(b1[k])) for k in range(nb))
    assert ranges[0][0] == 0 and ranges[-1][1] == lib.miso_net_grad_numel(m._handle)
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(nb - 1))          # contiguous, no overlap
    # completion order: the upper decoders' bucket comes first and lies between the encoders and the TCN in key order
    keys = [k for k, _ in m.named_parameters()]
    offs = np.cumsum([0] + [p.numel() for p in m.parameters()])
    first_dec = int(offs[keys.index("decoders.0.0.net.0.weight")])
    first_tcn = int(offs[[k.startswith("TCN.") for k in keys].index(True)])
    assert first_dec <= int(b0[0]) < int(b1[0]) == first_tcn
    st = torch.cuda.Stream()
    for k in range(nb):
        _lib.check(lib.miso_net_wait_grad_bucket(m._handle, k, ctypes.c_void_p(st.cuda_stream)), "miso_net_wait_grad_bucket")
    st.synchronize()
