"""Seeded synthetic weights with the reference's state_dict keys and shapes.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``) -- but deliberately free of
any oracle arithmetic: it is a key/shape table plus numpy PCG64 draws, so that the
reference (in the build container), the oracle and the CUDA path can all be loaded
with bit-identical parameters on any box, independent of torch's RNG and of module
construction order.

Key/shape table follows model.py:24-37 (ModuleList layout), model.py:401-433,
437-466 (conv wrappers, DenseBlock) and model.py:486-567 (TCN); SURVEY.md
section 8(b) lists the resulting 268 keys for the shipped config.
"""
import numpy as np
import torch


def param_shapes(cfg):
    """Ordered dict key -> shape for a NetConfig (oracle.miso_net_torch.NetConfig)."""
    shapes = {}
    nb = cfg.nb

    def conv(prefix, cin, cout):
        shapes[f"{prefix}.weight"] = (cout, cin, 3, 3)
        shapes[f"{prefix}.bias"] = (cout,)

    def deconv(prefix, cin, cout):
        shapes[f"{prefix}.weight"] = (cin, cout, 3, 3)
        shapes[f"{prefix}.bias"] = (cout,)

    def dense(prefix, c, g1, g2):
        for k in range(1, 6):
            conv(f"{prefix}.conv{k}.0", c + (k - 1) * g1, g1 if k < 5 else g2)

    for i in range(nb):
        cin, cout = cfg.en[i], cfg.en[i + 1]
        conv(f"encoders.{i}.0.conv2d" if i == 0 else f"encoders.{i}.0.net.0", cin, cout)
        if i < 5:
            dense(f"encoders.{i}.1", cout, cout, cout)
    for i in range(nb):
        cin, cout = 2 * cfg.de[i], cfg.de[i + 1]
        if i >= 2:
            dense(f"decoders.{i}.0", cin, cin // 2, cin)
            deconv(f"decoders.{i}.1.deconv2d" if i == nb - 1 else f"decoders.{i}.1.net.0", cin, cout)
        else:
            deconv(f"decoders.{i}.0.net.0", cin, cout)
    # self.decoders is registered before self.TCN (model.py:24-31)
    c = cfg.en[-1]
    for r in range(cfg.R):
        for x in range(cfg.X):
            for half in (2, 5):
                p = f"TCN.temporal_conv_net.{r}.{x}.net.{half}.net"
                shapes[f"{p}.0.weight"] = (c, 1, 3)
                shapes[f"{p}.1.weight"] = (1,)
                shapes[f"{p}.2.gamma"] = (1, c, 1)
                shapes[f"{p}.2.beta"] = (1, c, 1)
                shapes[f"{p}.3.weight"] = (c, c, 1)
    return shapes


def make_state_dict(cfg, seed=0):
    """PyTorch-default-like magnitudes (uniform +-1/sqrt(fan_in)) so activations stay
    O(1); gLN gamma/beta and the PReLU slope are perturbed away from their init values
    (1, 0, 0.25) so that those parameters are actually exercised."""
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in param_shapes(cfg).items():
        if key.endswith("gamma"):
            v = 1.0 + 0.2 * rng.standard_normal(shape)
        elif key.endswith("beta"):
            v = 0.1 * rng.standard_normal(shape)
        elif key.endswith(".1.weight") and len(shape) == 1:
            v = 0.25 + 0.1 * rng.random(shape)
        else:
            if len(shape) == 1:                     # bias: fan_in unknown here, keep small
                bound = 0.05
            elif "deconv2d" in key or (key.startswith("decoders") and ".net.0.weight" in key):
                bound = 1.0 / np.sqrt(shape[1] * 9)  # ConvTranspose2d: torch uses weight.size(1)*k*k
            else:
                bound = 1.0 / np.sqrt(np.prod(shape[1:]))
            v = rng.uniform(-bound, bound, shape)
        sd[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def state_dict_digest(sd):
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
