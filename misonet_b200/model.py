"""Host-side mirror of the reference's ``model.py`` interface for the hot path.

``MISO_1`` / ``MISO_3`` keep the reference constructor and ``forward`` signatures
(model.py:9, 76 and model.py:283, 350 of yuhogun0908/MISOnet) and the reference
``state_dict`` keys/shapes, so ``run.py:66-78``-style construction and
``load_state_dict(package['model_state_dict'])`` work unchanged.  The torch modules
below are *parameter containers only* (plumbing: storage, ``.cuda()``, ``state_dict``,
default initialisation in the reference's construction order); none of their
``forward`` methods is ever called.  The arithmetic runs in the CUDA library through
the C ABI (``include/misonet_b200.h``): pack (complex spectrogram -> bf16 hi/lo input planes) ->
``miso_net_forward`` -> unpack.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


# --------------------------------------------------------------------------- containers
class _Holder(nn.Module):
    """A module that only holds parameters; calling it is a bug."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("misonet_b200 parameter container: the computation runs in the CUDA library")


def _conv_unit(conv):
    # (conv, ELU, InstanceNorm2d) triple of model.py:411-414 / 428-431 / 443-445
    return nn.Sequential(conv, nn.ELU(), nn.InstanceNorm2d(conv.out_channels))


class _InitConv(_Holder):          # model.py:401-406
    def __init__(self, cin, cout):
        super().__init__()
        self.conv2d = nn.Conv2d(cin, cout, (3, 3), stride=(1, 1), padding=(1, 0))


class _Conv(_Holder):              # model.py:408-416
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.net = _conv_unit(nn.Conv2d(cin, cout, (3, 3), stride=stride, padding=(1, 0)))


class _LastDeconv(_Holder):        # model.py:418-423
    def __init__(self, cin, cout):
        super().__init__()
        self.deconv2d = nn.ConvTranspose2d(cin, cout, (3, 3), stride=(1, 1), padding=(1, 0))


class _Deconv(_Holder):            # model.py:425-433
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.net = _conv_unit(nn.ConvTranspose2d(cin, cout, (3, 3), stride=stride, padding=(1, 0)))


class _Dense(_Holder):             # model.py:437-466
    def __init__(self, c, g1, g2):
        super().__init__()
        for k in range(1, 6):
            setattr(self, f"conv{k}", _conv_unit(nn.Conv2d(c + (k - 1) * g1, g1 if k < 5 else g2, (3, 3), padding=(1, 1))))


class _GLN(_Holder):               # model.py:609-619
    def __init__(self, c):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(1, c, 1))
        self.beta = nn.Parameter(torch.zeros(1, c, 1))


class _DSConv(_Holder):            # model.py:553-561
    def __init__(self, c, dilation):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv1d(c, c, 3, stride=1, padding=dilation, dilation=dilation, groups=c, bias=False),
            nn.PReLU(), _GLN(c), nn.Conv1d(c, c, 1, bias=False))


class _TemporalBlock(_Holder):     # model.py:515-540
    def __init__(self, c, dilation):
        super().__init__()
        self.net = nn.Sequential(nn.InstanceNorm1d(c), nn.ELU(), _DSConv(c, dilation),
                                 nn.InstanceNorm1d(c), nn.ELU(), _DSConv(c, dilation))


class _TCN(_Holder):               # model.py:486-508
    def __init__(self, repeats, blocks, c):
        super().__init__()
        self.temporal_conv_net = nn.Sequential(
            *[nn.Sequential(*[_TemporalBlock(c, 2 ** x) for x in range(blocks)]) for _ in range(repeats)])


# --------------------------------------------------------------------------- training
class _NetFunction(torch.autograd.Function):
    """Autograd node of the network body for training (trainer.py:159-212: ``estimate = model(mix)``,
    ``loss.backward()``).  Forward = ``miso_net_forward_train`` (keeps the per-block TCN state in the training
    workspace), backward = ``miso_net_backward`` (include/misonet_b200.h); the parameter gradients come back as one
    flat buffer in the library's key order (= ``named_parameters()`` order) and are handed to autograd as views."""

    @staticmethod
    def forward(ctx, module, x_cl, B, T, F, *params):
        y_cl = module._run_body_train(x_cl, B, T, F)
        ctx.module, ctx.x_cl, ctx.shape, ctx.token = module, x_cl, (B, T, F), module._train_token
        ctx.param_shapes = [p.shape for p in params]
        return module._unpack(y_cl, B, T, F)

    @staticmethod
    def backward(ctx, gout):
        m = ctx.module
        if ctx.token != m._train_token:
            raise _lib.MisoError("the training workspace of this forward was overwritten by a later forward of the same "
                                 "module; call backward() before the next training forward")
        lib = _lib.load()
        B, T, F = ctx.shape
        S = m._out_ch // 2
        dev = ctx.x_cl.device
        g = gout.to(torch.complex64).contiguous()
        st = _lib.stream_ptr()
        with torch.cuda.device(dev):
            # persistent buffers: the backward's ~900 launches are replayed as a CUDA graph keyed by these pointers
            gy = m._buffer("gy_train", (B, T, F, (2 * S + 7) // 8 * 8), torch.float32, dev)
            _lib.check(lib.miso_grad_pack(_lib.ptr(g), _lib.ptr(gy), B, S, T, F, st), "miso_grad_pack")
            numel = lib.miso_net_grad_numel(m._handle)
            flat = m._buffer("flat_grads", (numel,), torch.float32, dev)
            ws = m._ws_train
            _lib.check(lib.miso_net_backward(m._handle, _lib.ptr(ctx.x_cl), _lib.ptr(gy), B, T, F, _lib.ptr(ws), ws.numel(),
                                             _lib.ptr(flat), st), "miso_net_backward")
            if m.data_parallel:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    m._allreduce_buckets(flat, numel, B, dev)
            # the step's gradients live in a copy, not in the buffer the next backward overwrites
            flat = flat.clone()
        # The 268 parameter gradients are handed over as views of that ONE buffer, set as ``p.grad`` directly: returned through
        # autograd they pass 268 AccumulateGrad nodes one by one (~0.5 ms of host time per step, measured 43.05 -> 42.54 ms).
        # A parameter that already holds a gradient (no ``zero_grad(set_to_none=True)``) accumulates.
        params = list(m._param_list)
        direct = len(params) == len(ctx.param_shapes) and all(p.requires_grad for p in params)
        grads, off = [], 0
        for i, shp in enumerate(ctx.param_shapes):
            n = int(torch.Size(shp).numel())
            g_i = flat[off:off + n].view(shp)
            off += n
            if direct:
                p = params[i]
                if p.grad is None:
                    p.grad = g_i
                else:
                    p.grad.add_(g_i)
                grads.append(None)
            else:
                grads.append(g_i)
        return (None, None, None, None, None, *grads)


# --------------------------------------------------------------------------- network
class _MisoNet(nn.Module):
    """Shared body of MISO_1 and MISO_3 (model.py:24-73 / 298-347)."""

    TCN_REPEATS, TCN_BLOCKS = 2, 7     # model.py:31

    def __init__(self, in_ch, out_ch, num_bottleneck, en_bottleneck_channels, de_bottleneck_channels, norm_type):
        super().__init__()
        if norm_type not in ("IN", None):
            raise ValueError(f"norm_type={norm_type!r}: only 'IN' (config/NN_BSS.yml:123) is implemented")
        # the reference mutates the caller's lists (model.py:16-17); we copy instead
        self._en = [int(c) for c in en_bottleneck_channels][:num_bottleneck]
        self._de = [int(c) for c in de_bottleneck_channels][:num_bottleneck]
        if len(self._en) != num_bottleneck or len(self._de) != num_bottleneck:
            raise ValueError("channel lists must have num_bottleneck entries")
        self.num_bottleneck = int(num_bottleneck)
        self._in_ch, self._out_ch = int(in_ch), int(out_ch)
        nb = self.num_bottleneck
        en = [self._in_ch] + self._en
        de = self._de + [self._out_ch]
        # construction order = the reference's (encoders, TCN, decoders) so that
        # torch.manual_seed(s) followed by construction gives the reference's weights,
        # registration order = the reference's (encoders, decoders, TCN) for state_dict.
        self.encoders = nn.ModuleList()
        self.decoders = nn.ModuleList()
        for i in range(nb):
            layers = []
            if i < 5:
                layers.append(_InitConv(en[i], en[i + 1]) if i == 0 else _Conv(en[i], en[i + 1], (1, 2)))
                layers.append(_Dense(en[i + 1], en[i + 1], en[i + 1]))
            elif i == nb - 1:
                layers.append(_Conv(en[i], en[i + 1], (1, 1)))
            else:
                layers.append(_Conv(en[i], en[i + 1], (1, 2)))
            self.encoders.append(nn.Sequential(*layers))
        # model.py:31 hard-wires 128; the documented 257-bin layout (model.py:30) needs the
        # bottleneck width, which is 128 for the shipped config.
        self.TCN = _TCN(self.TCN_REPEATS, self.TCN_BLOCKS, en[nb])
        for j in range(nb):
            cin, cout = 2 * de[j], de[j + 1]
            layers = []
            if j >= 2:
                layers.append(_Dense(cin, cin // 2, cin))
                layers.append(_LastDeconv(cin, cout) if j == nb - 1 else _Deconv(cin, cout, (1, 2)))
            elif j == 0:
                layers.append(_Deconv(cin, cout, (1, 1)))
            else:
                layers.append(_Deconv(cin, cout, (1, 2)))
            self.decoders.append(nn.Sequential(*layers))
        self.sigmoid = nn.Sigmoid()     # model.py:37 (unused there too)
        self._handle = None
        self._handle_device = None
        self._packed = {}
        self._ws = None
        self._ws_train = None     # training workspace: activations + statistics + gradient buffers of ONE forward
        self._train_token = 0
        self.data_parallel = False   # True: backward() all-reduces the parameter gradients across the ranks (training.py)
        self._bufs = {}           # persistent input-plane / output buffers: stable pointers keep the CUDA graph valid
        self._sync_tag = None
        self.use_graph = True     # replay the forward as a CUDA graph (include/misonet_b200.h, miso_net_set_graph)
        self.max_workspace_bytes = 48 << 30   # batches are processed in chunks that fit this
        # compute path of the stride-1 3x3 convs (include/misonet_b200.h, miso_net_set_mode):
        # "fp32" FMA | "bf16x3" tcgen05 split (fp32-grade) | "bf16" tcgen05 (throughput)
        # Default = the parity-grade tensor-core path bench.py measures; "fp32" is the 10x slower FMA reference path.
        self.conv_mode = "bf16x3"
        # eval-mode forwards never take the (activation-retaining, un-graphed) training path, even outside
        # torch.no_grad(); set True to differentiate through a module in eval() mode
        self.autograd_in_eval = False

    # ---- handle / weights ------------------------------------------------------------
    def _release(self):
        if self._handle is not None:
            _lib.load().miso_net_destroy(self._handle)
            self._handle = None
            self._packed = {}
            self._ws = None
            self._ws_train = None
            self._bufs = {}
            self._sync_tag = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _device(self):
        return next(self.parameters()).device

    def invalidate(self):
        """Forget what has been packed: the next forward re-uploads every parameter.  Needed after writes the change
        detection cannot see (``p.data.copy_()`` / ``.data.fill_()`` do not bump ``p._version``); ``load_state_dict`` and
        ``.to()`` / ``.cuda()`` call it themselves, and a replaced Parameter object is detected by identity."""
        self._packed = {}
        self._sync_tag = None
        self.__dict__.pop("_plist", None)
        self.__dict__.pop("_pnamed", None)

    def load_state_dict(self, *a, **k):
        self.invalidate()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def _ensure_handle(self):
        dev = self._device()
        if dev.type != "cuda":
            raise _lib.MisoError("misonet_b200 modules run on CUDA only (call .cuda() first); there is no CPU fallback")
        if self._handle is not None and self._handle_device == dev:
            return
        self._release()
        _lib.check_device(dev)
        lib = _lib.load()
        nb = self.num_bottleneck
        h = ctypes.c_void_p()
        en = (ctypes.c_int * nb)(*self._en)
        de = (ctypes.c_int * nb)(*self._de)
        with torch.cuda.device(dev):
            _lib.check(lib.miso_net_create(ctypes.byref(h), self._in_ch, self._out_ch, nb, en, de, self.TCN_REPEATS,
                                           self.TCN_BLOCKS), "miso_net_create")
        self._handle, self._handle_device = h, dev
        keys = [lib.miso_net_param_key(h, i).decode() for i in range(lib.miso_net_num_params(h))]
        mine = [k for k, _ in self.named_parameters()]
        if keys != mine:
            raise _lib.MisoError("parameter key table of the library differs from the module's state_dict")

    def _sync_params(self):
        """(Re)pack every parameter whose storage or version changed since the last call."""
        lib = _lib.load()
        st = _lib.stream_ptr()
        modes = {"fp32": 0, "bf16x3": 1, "bf16": 2}
        if self.conv_mode not in modes:
            raise ValueError(f"conv_mode must be one of {sorted(modes)}")
        _lib.check(lib.miso_net_set_mode(self._handle, modes[self.conv_mode]), "miso_net_set_mode")
        _lib.check(lib.miso_net_set_graph(self._handle, 1 if self.use_graph else 0), "miso_net_set_graph")
        # cheap change detection first: in-place updates bump _version, .cuda()/.to()/load_state_dict(assign) change storage
        params = self._param_list
        tag = (sum(p._version for p in params), sum(p.data_ptr() for p in params))
        if tag == self._sync_tag:
            return
        # the walk over the module tree is the expensive part of this function (268 parameters): when only versions moved
        # (an optimizer step) the cached key list is still valid; any storage change (.to(), load_state_dict(assign), a replaced
        # Parameter) re-walks
        named = self.__dict__.get("_pnamed")
        if named is None or self._sync_tag is None or tag[1] != self._sync_tag[1] or len(named) != len(params):
            named = list(self.named_parameters())
            if len(named) != len(params) or any(a is not b for (_, a), b in zip(named, params)):   # a Parameter was replaced
                self.__dict__["_plist"] = params = [p for _, p in named]
                tag = (sum(p._version for p in params), sum(p.data_ptr() for p in params))
            self.__dict__["_pnamed"] = named
        self._sync_tag = tag
        # every changed parameter in ONE library call (an optimizer step changes all 268: a call per parameter costs ~2 ms of
        # host time per training step); unchanged ones are passed as NULL
        ptrs = (ctypes.c_void_p * len(named))()
        keep, changed = [], 0
        for i, (key, p) in enumerate(named):
            tag = (p.data_ptr(), p._version)
            if self._packed.get(key) == tag:
                continue
            t = p.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
                keep.append(t)          # alive until the launch below has been enqueued (stream-ordered allocator)
            ptrs[i] = _lib.ptr(t)
            self._packed[key] = tag
            changed += 1
        if changed:
            _lib.check(lib.miso_net_set_params(self._handle, ptrs, len(named), st), "miso_net_set_params")

    @property
    def _param_list(self):
        pl = self.__dict__.get("_plist")
        if pl is None:
            pl = list(self.parameters())
            self.__dict__["_plist"] = pl
        return pl

    def _buffer(self, name, shape, dtype, dev):
        key = (name, tuple(shape), dtype, dev)
        t = self._bufs.get(key)
        if t is None:
            if len(self._bufs) > 64:
                self._bufs.clear()
            t = torch.empty(shape, dtype=dtype, device=dev)
            self._bufs[key] = t
        return t

    def _workspace(self, nbytes, dev):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self._ws

    def _ws_bytes(self, B, T, F):
        key = ("ws_bytes", B, T, F)
        v = self._bufs.get(key)
        if v is None:
            v = _lib.load().miso_net_workspace_bytes(self._handle, B, T, F)
            self._bufs[key] = v
        return v

    def _chunk(self, B, T, F):
        lib = _lib.load()
        if ("shape_ok", T, F) not in self._bufs:
            _lib.check(lib.miso_net_check_shape(self._handle, T, F), "miso_net_check_shape")
            self._bufs[("shape_ok", T, F)] = True
        per = self._ws_bytes(1, T, F)
        step = max(1, min(B, int(self.max_workspace_bytes // max(per, 1))))
        if 1 < step < B:
            # a sample's input planes are 64 * T * F bytes: an odd T * F makes every odd sample start 64-byte aligned
            # only, and miso_net_forward wants 128
            step -= step % 2
        return step

    def _input_planes(self, B, T, F, dev, train=False):
        """Input buffer in the library's plane layout (include/misonet_b200.h): uint8 [B, bytes per sample].
        A training forward gets its own buffer: the backward reads it (first layer's weight gradient), and an inference
        forward in between (validation) must not overwrite it."""
        per = _lib.load().miso_net_input_bytes(self._handle, 1, T, F)
        return self._buffer("x_train" if train else "x", (B, per), torch.uint8, dev)

    def _run_body(self, x_cl, B, T, F):
        """x_cl: input planes uint8 [B, bytes per sample] -> float32 [B,T,F,out_ch]."""
        lib = _lib.load()
        dev = x_cl.device
        y_cl = self._buffer("y", (B, T, F, self._out_ch), torch.float32, dev)
        step = self._chunk(B, T, F)
        nbytes = self._ws_bytes(step, T, F)
        ws = self._workspace(nbytes, dev)
        st = _lib.stream_ptr()
        for b0 in range(0, B, step):
            nbat = min(step, B - b0)
            _lib.check(lib.miso_net_forward(self._handle, _lib.ptr(x_cl[b0]), _lib.ptr(y_cl[b0]), nbat, T, F,
                                            _lib.ptr(ws), ws.numel(), st), "miso_net_forward")
        return y_cl

    def _training_pass(self):
        """True when this forward must be differentiable: autograd on, a parameter requires grad, and the module is in
        training mode (or ``autograd_in_eval``).  ``model.eval(); model(x)`` outside ``no_grad`` therefore stays on the
        graph-replayed inference path instead of silently allocating the training workspace."""
        return (torch.is_grad_enabled() and (self.training or self.autograd_in_eval)
                and any(p.requires_grad for p in self._param_list))

    def _run_body_train(self, x_cl, B, T, F):
        """Training forward: x_cl input planes -> float32 [B,T,F,out_ch]; leaves the workspace for the backward."""
        lib = _lib.load()
        dev = x_cl.device
        _lib.check(lib.miso_net_check_shape(self._handle, T, F), "miso_net_check_shape")
        nbytes = lib.miso_net_train_workspace_bytes(self._handle, B, T, F)
        if self._ws_train is None or self._ws_train.numel() < nbytes or self._ws_train.device != dev:
            self._ws_train = None
            self._ws_train = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        # persistent buffers (input planes, output, workspace): stable pointers let the library replay the training
        # forward as a CUDA graph from the third step on (miso_net_forward_train)
        y_cl = self._buffer("y_train", (B, T, F, self._out_ch), torch.float32, dev)
        self._train_token += 1
        _lib.check(lib.miso_net_forward_train(self._handle, _lib.ptr(x_cl), _lib.ptr(y_cl), B, T, F, _lib.ptr(self._ws_train),
                                              self._ws_train.numel(), _lib.stream_ptr()), "miso_net_forward_train")
        return y_cl

    def _allreduce_buckets(self, flat, numel, n_local, dev):
        """Data-parallel gradient reduction overlapped with the backward pass.  The flat gradient
        buffer is cut into the library's completion-ordered buckets (miso_net_grad_buckets: upper decoders, lower decoders,
        TCN, upper encoders, lower encoders); ``miso_net_backward`` -- already enqueued, still running -- records an event
        per bucket, and a communication stream all-reduces each bucket in place as soon as its event fires, so only the last
        (smallest) bucket's collective is exposed.  Every rank's gradient is that of the mean loss over ITS utterances
        (criterion.py:59), so buckets are weighted by n_local / sum(n_local): the result is the gradient of the mean over
        all utterances of the step.  Collective: NCCL all-reduce (SUM) of fp32, one call per bucket + one for the count."""
        import torch.distributed as dist
        lib = _lib.load()
        if self.__dict__.get("_comm_stream") is None or self._comm_stream.device != dev:
            self.__dict__["_comm_stream"] = torch.cuda.Stream(device=dev)
            cap = 8
            b0, b1 = (ctypes.c_int64 * cap)(), (ctypes.c_int64 * cap)()
            nb = lib.miso_net_grad_buckets(self._handle, b0, b1, cap)
            _lib.check(nb, "miso_net_grad_buckets")
            self.__dict__["_grad_buckets"] = [(int(b0[k]), int(b1[k])) for k in range(nb)]
        cs = self._comm_stream
        main = torch.cuda.current_stream(dev)
        flat.record_stream(cs)
        with torch.cuda.stream(cs):
            cnt = torch.full((1,), float(n_local), dtype=torch.float32, device=dev)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            w = float(n_local) / cnt                                   # device scalar: no host synchronisation
            for k, (lo, hi) in enumerate(self._grad_buckets):
                if hi <= lo:
                    continue
                _lib.check(lib.miso_net_wait_grad_bucket(self._handle, k, ctypes.c_void_p(cs.cuda_stream)), "miso_net_wait_grad_bucket")
                seg = flat[lo:hi]
                seg.mul_(w)
                dist.all_reduce(seg, op=dist.ReduceOp.SUM)
        main.wait_stream(cs)

    def _prepare(self, *tensors):
        self._ensure_handle()
        dev = self._handle_device
        out = []
        for t in tensors:
            _lib.require_cuda(t, "input")
            if t.device != dev:
                raise _lib.MisoError(f"input on {t.device}, module on {dev}")
            if not t.is_complex():
                raise TypeError("expected a complex spectrogram [B, Ch, T, F] (model.py:77-78 reads .real/.imag)")
            out.append(t.to(torch.complex64).contiguous())
        return out

    def _unpack(self, y_cl, B, T, F):
        S = self._out_ch // 2
        out = torch.empty(B, S, T, F, dtype=torch.complex64, device=y_cl.device)
        _lib.check(_lib.load().miso_unpack_complex(_lib.ptr(y_cl), _lib.ptr(out), B, S, T, F, _lib.stream_ptr()),
                   "miso_unpack_complex")
        return out

    def tap(self, name, B, T, F):
        """Parity/debug: an internal activation of the last forward as the reference sees it (NCHW)."""
        lib = _lib.load()
        cap = B * T * F * 6 * max(self._en + self._de)
        with torch.cuda.device(self._handle_device):
            buf = torch.empty(cap, dtype=torch.float32, device=self._handle_device)
            n = lib.miso_net_tap(self._handle, name.encode(), _lib.ptr(buf), cap, B, T, F, _lib.ptr(self._ws), _lib.stream_ptr())
        _lib.check(n, "miso_net_tap")
        return buf[:n]


class MISO_1(_MisoNet):
    """Separation network; same constructor and call signature as model.py:8-111."""

    def __init__(self, num_spks, num_ch, num_bottleneck, en_bottleneck_channels, de_bottleneck_channels, norm_type="IN"):
        super().__init__(2 * num_ch, 2 * num_spks, num_bottleneck, en_bottleneck_channels, de_bottleneck_channels, norm_type)
        self.num_spks, self.num_ch = int(num_spks), int(num_ch)

    def forward(self, mixture):
        """mixture: complex [B, Mic, T, F] -> complex64 [B, Spk, T, F] (model.py:76-111)."""
        return self.forward_shifts(mixture, (0,))

    def forward_shifts(self, mixture, shifts):
        """All circular microphone shifts of tester.py:1034,1049 as ONE batch:
        returns complex64 [len(shifts)*B, Spk, T, F]; row k*B+b is model(roll(mix,-shifts[k],1))[b]."""
        (mix,) = self._prepare(mixture)
        B, M, T, F = mix.shape
        if 2 * M != self._in_ch:
            raise ValueError(f"expected {self._in_ch // 2} microphones, got {M}")
        # every launch below goes to the MODULE's device and that device's current stream, whatever the caller's
        # current device is (the reference does model.cuda(gpu_num) without torch.cuda.set_device, run.py:68)
        with torch.cuda.device(self._handle_device):
            self._sync_params()
            n = len(shifts)
            train = self._training_pass()
            x_cl = self._input_planes(n * B, T, F, mix.device, train)
            arr = (ctypes.c_int * n)(*[int(s) for s in shifts])
            _lib.check(_lib.load().miso_pack_miso1(_lib.ptr(mix), _lib.ptr(x_cl), B, M, T, F, arr, n, _lib.stream_ptr()),
                       "miso_pack_miso1")
            if train:
                return _NetFunction.apply(self, x_cl, n * B, T, F, *self._param_list)
            y_cl = self._run_body(x_cl, n * B, T, F)
            return self._unpack(y_cl, n * B, T, F)


class MISO_3(_MisoNet):
    """Enhancement network; same constructor and call signature as model.py:282-395."""

    def __init__(self, num_spks, num_ch, num_bottleneck, en_bottleneck_channels, de_bottleneck_channels, norm_type="IN"):
        super().__init__(2 * (num_ch + 2), 2 * num_spks, num_bottleneck, en_bottleneck_channels, de_bottleneck_channels,
                         norm_type)
        self.num_spks, self.num_ch = int(num_spks), int(num_ch)

    def forward(self, mixture, MISO1, BF):
        """Positional order is what matters (model.py:350 names the arguments (mixture, MISO1, BF)
        but every caller passes (mix, beamformed, MISO1): tester.py:1242, trainer.py:398-414);
        the weights see channels [mix x M, second, third].
        mixture [B,M,T,F], second/third [B,1,T,F] complex -> complex64 [B, num_spks, T, F]."""
        mix, second, third = self._prepare(mixture, MISO1, BF)
        B, M, T, F = mix.shape
        if 2 * (M + 2) != self._in_ch:
            raise ValueError(f"expected {self._in_ch // 2 - 2} microphones, got {M}")
        if second.shape != (B, 1, T, F) or third.shape != (B, 1, T, F):
            raise ValueError("second/third inputs must be [B,1,T,F]")
        with torch.cuda.device(self._handle_device):     # see MISO_1.forward_shifts
            self._sync_params()
            train = self._training_pass()
            x_cl = self._input_planes(B, T, F, mix.device, train)
            _lib.check(_lib.load().miso_pack_miso3(_lib.ptr(mix), _lib.ptr(second), _lib.ptr(third), _lib.ptr(x_cl), B, M, T, F,
                                                   _lib.stream_ptr()), "miso_pack_miso3")
            if train:
                # trainer.py:398-414: the beamformed / MISO1 inputs are data (computed under no_grad or loaded from
                # disk, data.py:133-207), so only the parameters receive gradients
                return _NetFunction.apply(self, x_cl, B, T, F, *self._param_list)
            y_cl = self._run_body(x_cl, B, T, F)
            return self._unpack(y_cl, B, T, F)
