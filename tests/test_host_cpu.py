"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header
declares, host logic (shapes, sharding, key tables, error paths) and the world_size-2 gloo path."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

EN = [24, 32, 32, 32, 32, 64, 128]
DE = [128, 64, 32, 32, 32, 32, 24]


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "misonet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(miso_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from misonet_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/misonet_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"
    assert lib.miso_abi_version() == 1


def test_stft_frame_count_matches_oracle():
    from misonet_b200 import audio
    from oracle import miso_np
    for n in (1, 63, 64, 65, 1000, 1024, 12345, 32000, 64000):
        for nperseg, nover in ((256, 192), (512, 384)):
            assert audio.stft_num_frames(n, nperseg, nover) == miso_np.stft_num_frames(n, nperseg, nover)
    assert audio.stft_num_frames(32000) == 501


def test_shard_range_partitions():
    from misonet_b200.pipeline import shard_range
    for n in (0, 1, 7, 16, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_module_keys_shapes_and_reference_init():
    from misonet_b200.model import MISO_1, MISO_3
    from oracle import weights, ref_import
    from oracle import miso_net_torch as mnt
    en, de = list(EN), list(DE)
    m1 = MISO_1(2, 6, 7, en, de, "IN")
    assert en == EN and de == DE, "constructor must not mutate the caller's lists (model.py:16-17 does)"
    shapes = weights.param_shapes(mnt.NetConfig.miso1())
    sd = m1.state_dict()
    assert list(sd.keys()) == list(shapes.keys()) and len(sd) == 268
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    m3 = MISO_3(1, 6, 7, en, de, "IN")
    assert sum(p.numel() for p in m3.parameters()) == 2587382
    p8 = MISO_1(2, 6, 8, [24, 32, 32, 32, 32, 64, 128, 384], [384, 128, 64, 32, 32, 32, 32, 24], "IN")
    assert sum(p.numel() for p in p8.parameters()) == 8579704
    if ref_import.available():
        ns = ref_import.load()
        torch.manual_seed(0)
        r = ns.model.MISO_1(2, 6, 7, list(EN), list(DE), "IN")
        torch.manual_seed(0)
        o = MISO_1(2, 6, 7, list(EN), list(DE), "IN")
        rs, os_ = r.state_dict(), o.state_dict()
        assert list(rs.keys()) == list(os_.keys())
        assert all(torch.equal(rs[k], os_[k]) for k in rs), "same seed must give the reference's initial weights"


def test_no_cpu_fallback():
    from misonet_b200 import _lib, audio, beamforming, criterion
    from misonet_b200.model import MISO_1
    m = MISO_1(2, 6, 7, list(EN), list(DE), "IN")
    x = torch.zeros(1, 6, 4, 129, dtype=torch.complex64)
    with torch.no_grad():
        with pytest.raises(_lib.MisoError):
            m(x)
    with pytest.raises(_lib.MisoError):
        audio.stft(torch.zeros(1000, 2))
    with pytest.raises(_lib.MisoError):
        criterion.loss_Enhance(x, x)
    with pytest.raises(_lib.MisoError):
        beamforming.mvdr(torch.zeros(1, 1, 6, 4, 9, dtype=torch.complex64), torch.zeros(1, 6, 4, 9, dtype=torch.complex64))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under misonet_b200/ may reference it."""
    pkg = os.path.join(ROOT, "misonet_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), fn
                assert "/root/reference" not in text, fn


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from misonet_b200 import distributed as D
rank, world, local = D.init_from_env("gloo")
n_total = 7
lo, hi = D.shard_range(n_total, rank, world)
# a per-utterance "loss" and "decision" that only depend on the global utterance index
loss = sum(float(i + 1) for i in range(lo, hi))
idx = torch.tensor([i % 2 for i in range(lo, hi)], dtype=torch.long)
mean, count = D.reduce_metrics(loss, hi - lo)
allidx = D.gather_perm_indices(idx, n_total)
mx = D.max_over_ranks(10.0 * (rank + 1))
from misonet_b200 import continuous as C
wave = torch.stack([torch.full((2, 5), float(i)) for i in range(lo, hi)]) if hi > lo else torch.zeros(0, 2, 5)
allwave = C.gather_chunks(wave, n_total)          # the long-recording path: chunk waveforms gathered in chunk order
assert allwave.shape == (n_total, 2, 5) and allwave[:, 0, 0].tolist() == [float(i) for i in range(n_total)]
# gradient all-reduce of the training step: rank r holds the mean gradient of its (hi - lo) utterances
params = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
per_utt = lambda i, p: torch.full_like(p, float(i + 1))
for p in params:
    p.grad = sum(per_utt(i, p) for i in range(lo, hi)) / max(hi - lo, 1)
tot = D.allreduce_gradients(params, n_local=hi - lo)
assert tot == n_total
for p in params:
    assert torch.allclose(p.grad, torch.full_like(p, 4.0)), p.grad      # mean over all 7 utterances of (i + 1)
# training._backward on a module without the library's in-backward reduction: mean-loss gradients of unequal shards
# must come out as the gradient of the mean over ALL utterances
from misonet_b200 import training as TR
torch.manual_seed(0)
lin = torch.nn.Linear(4, 1)
xs = torch.arange(n_total * 4, dtype=torch.float32).reshape(n_total, 4) / 10.0
full = lin(xs).pow(2).mean()
gfull = torch.autograd.grad(full, list(lin.parameters()))
TR._backward(lin, lin(xs[lo:hi]).pow(2).mean(), hi - lo)
for p, gexp in zip(lin.parameters(), gfull):
    assert torch.allclose(p.grad, gexp, rtol=1e-5, atol=1e-6), (p.grad, gexp)
D.barrier()
assert count == n_total and abs(mean - 4.0) < 1e-12, (mean, count)
assert allidx.tolist() == [i % 2 for i in range(n_total)], allidx
assert mx == 10.0 * world
if rank == 0:
    print("GLOO_OK", world)
dist.destroy_process_group()
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GLOO_OK 2" in res.stdout


def test_chunk_signal_follows_the_reference_rule():
    """dataloader/data.py:538-597: full chunks, then the remainder zero-padded by gap."""
    from misonet_b200 import continuous as C
    wav = torch.arange(25, dtype=torch.float32).reshape(25, 1).repeat(1, 2)
    chunks, gap = C.chunk_signal(wav, 10)
    assert chunks.shape == (3, 10, 2) and gap == 5
    assert chunks[2, :5, 0].tolist() == [20., 21., 22., 23., 24.] and float(chunks[2, 5:].abs().sum()) == 0.0
    chunks, gap = C.chunk_signal(wav[:20], 10)
    assert chunks.shape == (2, 10, 2) and gap == 0
    chunks, gap = C.chunk_signal(wav[:7], 10)         # shorter than one chunk (data.py:538-541)
    assert chunks.shape == (1, 10, 2) and gap == 3


def test_wav_container_pcm24_round_trip(tmp_path):
    """The reference writes int16 samples with sf.write(..., 'PCM_24') (tester.py:447, 971-972): a 24-bit container holding
    the int16 value shifted left by 8.  Checked with the standard library's wave reader."""
    import wave
    import numpy as np
    from misonet_b200 import audio
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32767, size=(1000, 2), dtype=np.int16)
    path = str(tmp_path / "x.wav")
    audio.write_wav(path, pcm, 8000)
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 3, 8000, 1000)
        raw = np.frombuffer(w.readframes(1000), dtype=np.uint8).reshape(1000, 2, 3).astype(np.int32)
    v = raw[..., 0] | (raw[..., 1] << 8) | (raw[..., 2] << 16)
    v = np.where(v >= 1 << 23, v - (1 << 24), v)
    assert np.array_equal(v, pcm.astype(np.int32) << 8)
    x, fs = audio.read_wav(path)
    assert fs == 8000 and x.shape == (1000, 2)
    assert np.array_equal(x, (pcm.astype(np.float32) / 32768.0))          # sf.read's normalisation
    audio.write_wav(path, pcm[:, 0], 16000, subtype="PCM_16")
    x, fs = audio.read_wav(path)
    assert fs == 16000 and np.array_equal(x[:, 0], pcm[:, 0].astype(np.float32) / 32768.0)


def test_dropin_mixes_into_the_real_tester():
    """INTEGRATION.md's `class Tester(B200HotPath, tester.Tester_Enhance)`: when the reference is present (build
    container only) the mix-in must resolve our methods ahead of the reference's, and every overridden name must
    exist on the reference class (tester.py:979 ISTFT, :992 STFT, :1014 MISO1_Inference, :1071 Apply_Beamforming,
    :1231 MISO3_inference) -- a renamed reference method would otherwise silently bypass the drop-in."""
    import inspect
    from misonet_b200 import dropin
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    if not ref_import.available():
        import pytest
        pytest.skip("/root/reference is not present (GPU box)")
    ref = ref_import.load().tester.Tester_Enhance
    ours = [n for n, f in vars(dropin.B200HotPath).items() if inspect.isfunction(f) and not n.startswith("_")]
    assert sorted(ours) == sorted(["ISTFT", "STFT", "MISO1_Inference", "Apply_Beamforming", "MISO3_inference"])

    class T(dropin.B200HotPath, ref):
        pass

    for name in ours:
        assert hasattr(ref, name), f"reference Tester_Enhance has no method {name}"
        assert getattr(T, name) is getattr(dropin.B200HotPath, name), f"MRO does not resolve {name} to the drop-in"
        # same positional parameters (names and order), so the reference's own call sites bind unchanged
        want = list(inspect.signature(getattr(ref, name)).parameters)
        got = list(inspect.signature(getattr(dropin.B200HotPath, name)).parameters)
        assert got[:len(want)] == want, (name, got, want)
    assert T.inference is ref.inference            # the reference's loop itself is inherited untouched
    t = T.__new__(T)
    t.device = 0
    assert str(t._b200_device()) == "cuda:0"


def test_model_defaults_and_invalidate():
    """conv_mode defaults to the benched parity-grade path; invalidate() / load_state_dict / .to() reset the packed-weight
    cache; an eval-mode forward never selects the training path (host logic only: no CUDA call is made here)."""
    from misonet_b200.model import MISO_1
    m = MISO_1(2, 6, 7, [24, 32, 32, 32, 32, 64, 128], [128, 64, 32, 32, 32, 32, 24], "IN")
    assert m.conv_mode == "bf16x3"
    m._packed["x"] = 1
    m._sync_tag = (1, 2)
    m.invalidate()
    assert m._packed == {} and m._sync_tag is None
    m._packed["x"] = 1
    m.load_state_dict(m.state_dict())
    assert m._packed == {}
    m._packed["x"] = 1
    m.float()
    assert m._packed == {}
    m.eval()
    assert not m._training_pass()
    m.train()
    assert m._training_pass()
    with torch.no_grad():
        assert not m._training_pass()
    m.eval()
    m.autograd_in_eval = True
    assert m._training_pass()


def test_documented_switches_exist_in_the_sources():
    """Every MISO_* environment switch that INTEGRATION.md / DESIGN.md name is read somewhere in the sources (stale docs
    after an experiment is removed are the failure this guards against)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = ""
    for d, _, files in os.walk(os.path.join(root, "misonet_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py")):
                src += open(os.path.join(d, f), errors="ignore").read()
    read = set(re.findall(r'getenv\("(MISO_[A-Z0-9_]+)"\)', src)) | set(re.findall(r'environ[^\n]*"(MISO_[A-Z0-9_]+)"', src))
    removed = {"MISO_RS_FUSE", "MISO_RS_GROUPKERNEL", "MISO_RS_FORK", "MISO_RS_DBG"}   # named in DESIGN.md as removed experiments
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        text = open(os.path.join(root, doc)).read()
        for name in set(re.findall(r"`(MISO_[A-Z0-9_]+)(?:=[^`]*)?`", text)):
            if (name.startswith("MISO_E_") or name.startswith("MISO_PROF") or name in removed or name == "MISO_OK"
                    or re.fullmatch(r"MISO_\d", name)):     # MISO_1 / MISO_3 are classes of the reference
                continue
            assert name in read, f"{doc} names {name}, which no source file reads"
