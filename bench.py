#!/usr/bin/env python
"""bench.py -- frames/sec of the MISO hot path on B200 (see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_gpu] [--workload NAME]

Workloads (BASELINE.json configs):
  miso1_paper   (default, configs[1]) MISO_1 separation forward, per-GPU batch 16 x 6 mics x 257 bins x
                500 frames, 8-block "paper" layout (model.py:13-14,30 comments)
  miso1_ref     the same on the shipped 7-block / 129-bin / 501-frame layout
  pipeline_ref  (configs[2]) STFT -> MISO1 x6 shifts -> align -> MVDR x2 -> MISO3 x2, per-GPU batch 32, REF layout
  pipeline_paper  the same at the PAPER shape (8 blocks, 512-point STFT: 257 bins x 500 frames), per-GPU batch 32
  train_paper   (configs[3] per-GPU shape) MISO_1 training step: forward + loss_uPIT + backward + gradient all-reduce +
                Adam, 8 utterances per GPU, PAPER layout (tools/train_step.py prints the line)

One "step" = one pass of the workload over one synthetic batch per GPU.  `value` is whole-job
frames/s with inputs resident in HBM; `e2e` is the same through the public API with pinned host
buffers (H2D of the inputs and D2H of the result inside the timed region).  Under torchrun every
rank processes its own batch (weak scaling, no data-path collective); time = max over ranks.

--impl reference times the reference's own CPU algorithm for the same workload (the full per-GPU batch per step)
on the host cores.  /root/reference does not exist on the GPU box, so this is the oracle port
(oracle/miso_net_torch.py, pinned to the real reference by tests/golden); rank 0 only.
--impl torch_gpu is a CONTEXT line, not the reference arm: the same port run unchanged on the B200 through
stock PyTorch / cuDNN (fp32 with TF32 off and on) -- the library incumbent on the same chip; the default line
carries it as `gpu_library_baseline`.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...") and debug output go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

LAYOUTS = {
    "REF": ([24, 32, 32, 32, 32, 64, 128], [128, 64, 32, 32, 32, 32, 24]),
    "PAPER": ([24, 32, 32, 32, 32, 64, 128, 384], [384, 128, 64, 32, 32, 32, 32, 24]),
}
WORKLOADS = {
    "miso1_paper": dict(layout="PAPER", B=16, M=6, T=500, F=257, kind="miso1",
                        desc="MISO1 separation fwd, batch 16 x 6ch x 257bin x 500fr per GPU (BASELINE configs[1])"),
    "miso1_ref": dict(layout="REF", B=16, M=6, T=501, F=129, kind="miso1",
                      desc="MISO1 separation fwd, batch 16 x 6ch x 129bin x 501fr per GPU (shipped config shape)"),
    "pipeline_ref": dict(layout="REF", B=32, M=6, T=501, F=129, kind="pipeline", n_samples=32000, nperseg=256, noverlap=192,
                         desc="STFT->MISO1x6->align->MVDRx2->MISO3x2, batch 32 per GPU (BASELINE configs[2], REF shape)"),
    "pipeline_paper": dict(layout="PAPER", B=32, M=6, T=500, F=257, kind="pipeline", n_samples=499 * 128, nperseg=512, noverlap=384,
                           desc="STFT->MISO1x6->align->MVDRx2->MISO3x2, batch 32 x 6ch x 257bin x 500fr per GPU (BASELINE configs[2], PAPER shape)"),
}
PROF_FAMILIES = {0: "conv_fp32_kernel (fp32 FMA implicit-GEMM conv / deconv / pointwise)",
                 1: "conv_tc_kernel (tcgen05 implicit-GEMM 3x3 (de)conv, shifted-descriptor im2col: strided / transposed / narrow stages)",
                 2: "tcn_pw_kernel (tcgen05 pointwise convs of the TCN)",
                 3: "mvdr kernels (scm -> eig6 -> solve -> apply per call: HBM-bound streaming + latency-bound 6x6 eigen/solve)",
                 5: "conv_*_prep_kernel (per-sample weight images and border-bias sums of the tensor-core convs)",
                 4: "conv_rs_kernel (row-streaming tcgen05 3x3 conv, frame taps merged into N: the DenseBlock convs)"}
CONV_MODES = {"fp32": ("f32", "fp32 FMA everywhere (reference-grade, ~2e-6 rel. error)"),
              "bf16x3": ("bf16x3", "tcgen05 bf16 hi/lo split, 3 MMAs per product, fp32 accumulate (parity-grade, ~2e-5 rel. error)"),
              "bf16": ("bf16", "tcgen05 bf16 operands, fp32 accumulate (throughput mode, ~1e-2 rel. error: outside north_star's 1e-3)")}
# algorithmic conv-stack work per utterance, SURVEY.md section 8(d) / appendix A
GFLOP_PER_UTT = {"REF": 75.151, "PAPER": 163.37}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                        source="measured (MEASURED_PEAKS.json, sustained bf16)")
        except Exception:
            pass
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


def rand_spec(seed, shape, device):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(*shape, 2, generator=g, dtype=torch.float32)
    return torch.view_as_complex(x).contiguous().to(device) if device != "pinned" else torch.view_as_complex(x).contiguous().pin_memory()


def oracle_state_dict(kind, layout, seed):
    from oracle import weights
    from oracle import miso_net_torch as mnt
    cfg = mnt.NetConfig.miso1(layout=layout) if kind == "miso1" else mnt.NetConfig.miso3(layout=layout)
    return cfg, weights.make_state_dict(cfg, seed)


def make_state_dict_np(model, seed):
    """Seeded random weights with PyTorch-default magnitudes, generated without touching oracle/."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if k.endswith("gamma"):
            sd[k] = torch.ones_like(v)
        elif k.endswith("beta"):
            sd[k] = torch.zeros_like(v)
        elif v.numel() == 1:
            sd[k] = torch.full_like(v, 0.25)
        else:
            fan = v[0].numel() if v.dim() > 1 else 64
            sd[k] = (torch.rand(v.shape, generator=g) * 2 - 1) / np.sqrt(fan)
    return sd


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower() == "active":
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------ ours
def run_ours(args, wl, rank, world, local):
    from misonet_b200 import _lib, distributed as D
    from misonet_b200.model import MISO_1, MISO_3
    from misonet_b200 import pipeline
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    en, de = LAYOUTS[wl["layout"]]
    B, M, T, F = wl["B"], wl["M"], wl["T"], wl["F"]
    m1 = MISO_1(2, M, len(en), list(en), list(de), "IN")
    m1.load_state_dict(make_state_dict_np(m1, 0))
    m1 = m1.cuda(dev).eval()
    m1.conv_mode = args.conv_mode
    if wl["kind"] == "pipeline":
        m3 = MISO_3(1, M, len(en), list(en), list(de), "IN")
        m3.load_state_dict(make_state_dict_np(m3, 1))
        m3 = m3.cuda(dev).eval()
        m3.conv_mode = args.conv_mode
        pipe = pipeline.MisoBfMiso(m1, m3, nperseg=wl["nperseg"], noverlap=wl["noverlap"])
        g = torch.Generator().manual_seed(100 + rank)
        host_in = (0.05 * torch.randn(B, wl["n_samples"], M, generator=g)).pin_memory()
        dev_in = host_in.to(dev)

        def step(x):
            return pipe(x)["enhanced"]
    else:
        host_in = rand_spec(100 + rank, (B, M, T, F), "pinned")
        dev_in = host_in.to(dev)

        def step(x):
            return m1(x)

    frames_per_step = B * T
    with torch.no_grad():
        out = step(dev_in)
        out0 = out[0:1].cpu()
        host_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        for _ in range(max(args.warmup, 3) - 1):
            step(dev_in)
        torch.cuda.synchronize()

        # ---- resident-input throughput (value): CUDA-graph replays, no per-launch instrumentation ----
        sampler = ClockSampler(local) if rank == 0 else None
        import ctypes
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        torch.cuda.synchronize()
        launches0 = _lib.launch_count()
        e0.record()
        for _ in range(args.steps):
            step(dev_in)
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - launches0
        clocks = sampler.stop() if sampler else None

        # ---- roofline pass: the same steps with CUDA-event nodes around every conv launch (baked into the
        # replayed graph), collected after each step ----
        fams = {fname: dict(ms=0.0, flops=0.0, bytes=0.0, exec=0.0, launches=0) for fname in PROF_FAMILIES.values()}
        lib.miso_prof_enable(1)
        step(dev_in)                       # captures the instrumented graph
        torch.cuda.synchronize()
        for fid in PROF_FAMILIES:
            lib.miso_prof_collect(fid, None, None, None, None)
        e0.record()
        for _ in range(args.steps):
            step(dev_in)
            torch.cuda.synchronize()
            for fid, fname in PROF_FAMILIES.items():
                cm, cf, cb, ce, cn = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_uint64()
                _lib.check(lib.miso_prof_collect2(fid, ctypes.byref(cm), ctypes.byref(cf), ctypes.byref(cb), ctypes.byref(ce), ctypes.byref(cn)))
                f = fams[fname]
                f["ms"] += cm.value
                f["flops"] += cf.value
                f["bytes"] += cb.value
                f["exec"] += ce.value
                f["launches"] += int(cn.value)
        e1.record()
        torch.cuda.synchronize()
        lib.miso_prof_enable(0)

        # ---- end to end through the public API: every step copies its input from pinned host memory (H2D) and
        # its result back (D2H) inside the timed region.  Copies run on a second stream and overlap the previous /
        # next step's compute through two device input buffers, as a streaming caller would do. ----
        copy_s = torch.cuda.Stream(device=dev)      # H2D
        back_s = torch.cuda.Stream(device=dev)      # D2H (its own stream, or it would serialise behind the next H2D)
        main_s = torch.cuda.current_stream(dev)
        dev_bufs = [torch.empty_like(dev_in) for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]

        def e2e_loop(nsteps):
            for i in range(nsteps):
                k = i & 1
                with torch.cuda.stream(copy_s):
                    if i >= 2:
                        copy_s.wait_event(ev_free[k])          # step i-2 finished reading this buffer
                    dev_bufs[k].copy_(host_in, non_blocking=True)
                    ev_in[k].record(copy_s)
                main_s.wait_event(ev_in[k])
                out_i = step(dev_bufs[k])
                ev_free[k].record(main_s)
                out_i.record_stream(back_s)
                with torch.cuda.stream(back_s):
                    back_s.wait_event(ev_free[k])
                    host_out.copy_(out_i, non_blocking=True)
            copy_s.synchronize()
            back_s.synchronize()

        e2e_loop(2)
        torch.cuda.synchronize()
        D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        torch.cuda.synchronize()
        ms_e2e_wall = (time.perf_counter() - t0) * 1e3
        D.barrier()
        ms_e2e = max(e0.elapsed_time(e1), ms_e2e_wall)   # the D2H tail runs on the copy stream: take the host clock too

    ms = D.max_over_ranks(ms, dev)
    ms_e2e = D.max_over_ranks(ms_e2e, dev)
    if rank != 0:
        return None
    pk = peaks()
    total_frames = frames_per_step * world * args.steps
    dom = max(fams, key=lambda k: fams[k]["ms"])
    conv_ms, conv_flops, conv_launches = fams[dom]["ms"], fams[dom]["flops"], fams[dom]["launches"]
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tp = next((os.path.join(ROOT, "profiles", f) for f in ("r2_ncu_traffic_conv_rs_bf16x3.json", "r1_c_ncu_traffic_conv_rs_bf16x3.json")
               if os.path.isfile(os.path.join(ROOT, "profiles", f))), None)
    if args.workload == "miso1_paper" and args.conv_mode == "bf16x3" and args.batch == 0 and dom.startswith("conv_rs") and tp:
        try:
            import hashlib
            tj = json.load(open(tp))
            traffic = float(tj["family_dram_bytes_per_launch"])
            cur = hashlib.sha1(open(os.path.join(ROOT, "misonet_b200", "csrc", "conv_rs.cu"), "rb").read()).hexdigest()
            stamp = tj.get("kernel_source_sha1")
            fresh = "capture taken from this very conv_rs.cu" if stamp == cur else (
                "STALE: conv_rs.cu changed since the capture" if stamp else "capture predates the source stamp (round 1)")
            traffic_src = (f"profiles/{os.path.basename(tp)} (ncu dram__bytes_read.sum + dram__bytes_write.sum of this workload's conv_rs + prep "
                           f"launches, per conv launch; ncu cannot run inside the timed bench; {fresh})")
        except Exception:
            pass
    tc = [v for k, v in fams.items() if k.split()[0] in ("conv_tc_kernel", "conv_rs_kernel", "tcn_pw_kernel")]
    tc_ms, tc_fl = sum(v["ms"] for v in tc), sum(v["flops"] for v in tc)
    all_tc = {"ms_per_step": tc_ms / args.steps, "tflops": tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0,
              "frac": tc_fl / (tc_ms * 1e-3) / 1e12 / pk["bf16_tflops"] if tc_ms > 0 else 0.0, "share_of_step": tc_ms / ms if ms > 0 else None}
    fam_report = {k: {"ms_per_step": v["ms"] / args.steps, "share_of_step": v["ms"] / ms if ms > 0 else None,
                      "launches_per_step": v["launches"] // max(args.steps, 1),
                      "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0,
                      "executed_tflops": v["exec"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0,
                      "algorithmic_gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0}
                  for k, v in fams.items() if v["launches"]}
    line = {
        "metric": "frames/sec MISO-BF-MISO fwd 6ch/257bin at 1/2/4/8 GPU; SI-SDR vs ref",
        "value": total_frames / (ms * 1e-3),
        "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": CONV_MODES[args.conv_mode][0], "data": "synthetic (seeded random spectrograms / waveforms, seeded random weights)",
        "config": {"workload": wl["desc"], "layout": wl["layout"], "per_gpu_batch": B, "global_batch": B * world,
                   "frames": T, "bins": F, "mics": M, "parallelism": f"utterance-sharded x{world}, no data-path collective",
                   "conv_mode": f"{args.conv_mode}: {CONV_MODES[args.conv_mode][1]}",
                   "l2": "activation working set per step is several GB (>> 126 MB L2), so every step starts cold"},
        "e2e": {"value": total_frames / (ms_e2e * 1e-3), "unit": "frames/s",
                "h2d_bytes_per_step": int(host_in.numel() * host_in.element_size()),
                "d2h_bytes_per_step": int(host_out.numel() * host_out.element_size())},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": dom,
                     "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_tflops"],
                     "executed_frac": (fams[dom]["exec"] / (conv_ms * 1e-3) / 1e12 / pk["bf16_tflops"]) if conv_ms > 0 else None,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": fams[dom]["bytes"] / max(conv_launches, 1), "peak_source": pk["source"],
                     "launches": int(conv_launches), "kernel_ms_per_step": conv_ms / args.steps,
                     "share_of_step": conv_ms / ms if ms > 0 else None,
                     "families": fam_report,
                     "all_tensor_core_kernels": all_tc,
                     "note": "algorithmic 2*MAC of the launches of the dominant kernel / summed "
                             "CUDA-event time of those launches (event-record nodes around every conv launch of the replayed CUDA "
                             "graph, a second pass of the same steps right after the timed one; the per-sample operand preparation launches are their "
                             "own entry under families).  bf16x3 spends 3 MMAs per "
                             "algorithmic product and an SS-mode MMA is shared-memory-operand bound at (4096 + 32 N) / 128 cycles, "
                             "so the kernel's own ceiling is 0.29 of the bf16 peak at N = 3 cout = 96 (0.86 in bf16 mode); executed_frac counts "
                             "the MMAs actually issued (3 per product, N padded to 3 x 32 for the cout-24 layers, M = 128 rows for 127 bins, "
                             "K padded to 16 channels) against the same peak: how busy the tensor pipe is; "
                             "all_tensor_core_kernels aggregates every tcgen05 kernel of the step"},
    }
    mv = fams.get(PROF_FAMILIES[3])
    if mv and mv["launches"]:
        gbs = mv["bytes"] / (mv["ms"] * 1e-3) / 1e9
        line["roofline"]["mvdr"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                                    "ms_per_step": mv["ms"] / args.steps, "calls_per_step": mv["launches"] // max(args.steps, 1),
                                    "note": "algorithmic bytes (mixture + sources in, outputs out: 160 T F per utterance) / CUDA-event time of "
                                            "the four kernels of each miso_mvdr_fwd call; the 6x6 eigenvector and solve kernels are latency bound"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl, steps=3)
        if wl["kind"] == "miso1":   # "SI-SDR vs ref" of the metric: utterance 0 of the bench batch against the oracle (checker only)
            line["parity"] = parity_vs_oracle(wl, m1, host_in, out0)
            line["gpu_library_baseline"] = torch_gpu_baseline(wl, dev, steps=3, value_ours=line["value"])
    return line


def parity_vs_oracle(wl, model, host_in, out0):
    """Part of the cpu_baseline leg: the oracle (reference algorithm, fp32 CPU) on utterance 0 of the bench batch with the
    bench model's weights, against the CUDA result of the timed configuration."""
    from oracle import miso_net_torch as mnt
    cfg = mnt.NetConfig.miso1(layout=wl["layout"])
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = mnt.miso1_forward(sd, cfg, host_in[0:1].clone())
    err = float((out0 - ref).norm() / ref.norm())
    return {"rel_err": err, "si_sdr_vs_ref_db": -20.0 * float(np.log10(max(err, 1e-30))), "tolerance": 1e-3,
            "sample": "utterance 0 of the bench batch, complex64 [1,2,T,F], oracle = reference algorithm in fp32 on the CPU"}


# ------------------------------------------------------------------------------------ library incumbent on the GPU
def torch_gpu_baseline(wl, dev, steps=3, value_ours=None, batch=None):
    """Context for the headline ratio: the reference algorithm (oracle port = the same ATen ops as
    model.py) run UNCHANGED on the same B200 through stock PyTorch / cuDNN, fp32 with TF32 off and on, same shape and batch.
    A baseline leg like cpu_baseline: the only place the bench runs oracle/ on the GPU."""
    from oracle import miso_net_torch as mnt
    layout, M, T, F = wl["layout"], wl["M"], wl["T"], wl["F"]
    B = batch or wl["B"]
    cfg, sd = oracle_state_dict("miso1", layout, 0)
    sd_cpu = sd
    sd = {k: v.to(dev) for k, v in sd.items()}
    mix = rand_spec(100, (B, M, T, F), "cpu")
    mix_d = mix.to(dev)
    ref1 = mnt.miso1_forward(sd_cpu, cfg, mix[0:1])                    # fp32 CPU = the parity yardstick
    out = {}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            y = None
            for _ in range(2):
                y = mnt.miso1_forward(sd, cfg, mix_d)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                y = mnt.miso1_forward(sd, cfg, mix_d)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            err = float((y[0:1].cpu() - ref1).norm() / ref1.norm())
            out[name] = {"value": B * T / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "rel_err_vs_fp32_cpu": err}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["what"] = (f"oracle/miso_net_torch.miso1_forward on cuda through stock PyTorch {torch.__version__} / cuDNN, batch {B}, "
                   f"{layout} layout, {T} frames x {F} bins, inputs resident, cudnn.benchmark on, {steps} steps after 2 warm-ups")
    if value_ours:
        out["ours_over_fp32"] = value_ours / out["fp32"]["value"]
        out["ours_over_tf32"] = value_ours / out["tf32"]["value"]
    return out


def run_torch_gpu(args, wl, rank, world, local):
    if rank != 0:
        return None
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    g = torch_gpu_baseline(wl, dev, steps=max(1, args.steps))
    return {"impl": "torch_gpu", "metric": "frames/sec MISO-BF-MISO fwd 6ch/257bin at 1/2/4/8 GPU; SI-SDR vs ref",
            "value": g["fp32"]["value"], "unit": "frames/s", "n_gpus": 1, "steps": max(1, args.steps), "warmup": 2,
            "ms_per_step": g["fp32"]["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded random spectrograms, seeded random weights)",
            "config": {"workload": wl["desc"], "layout": wl["layout"], "per_gpu_batch": wl["B"], "frames": wl["T"], "bins": wl["F"],
                       "mics": wl["M"], "note": "context line (library incumbent on the same GPU), not the reference arm"},
            "gpu_library_baseline": g}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_baseline(wl, steps=1, warmup=1, utts=1):
    """The reference's CPU algorithm (oracle port) for the same workload on a bounded sample."""
    from oracle import miso_net_torch as mnt
    from oracle import miso_np
    torch.set_num_threads(os.cpu_count() or 1)
    layout, M, T, F = wl["layout"], wl["M"], wl["T"], wl["F"]
    cfg1, sd1 = oracle_state_dict("miso1", layout, 0)
    mix = rand_spec(7, (utts, M, T, F), "cpu")
    if wl["kind"] == "pipeline":
        cfg3, sd3 = oracle_state_dict("miso3", layout, 1)

        def step():
            m1, _ = mnt.miso1_inference(sd1, cfg1, mix, 0)
            for s in range(2):
                src = m1[s].permute(0, 3, 1, 2).numpy()
                bf = miso_np.apply_beamforming(src, mix.permute(0, 3, 1, 2).numpy())
                mnt.miso3_forward(sd3, cfg3, mix, torch.from_numpy(bf).unsqueeze(1), m1[s][:, 0:1])
    else:
        def step():
            mnt.miso1_forward(sd1, cfg1, mix)
    for _ in range(warmup):
        step()
    best = float("inf")
    t_all0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t0)
    total = time.perf_counter() - t_all0
    return {"value": utts * T / best, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{utts} utterance(s) of the same workload ({wl['layout']} layout, {T} frames x {F} bins), "
                      f"fp32 torch CPU ops, best of {max(steps, 1)} after {warmup} warm-up",
            "seconds_best": best, "seconds_total": total}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return None
    steps = max(1, min(args.steps, 5))
    # the arm's config is the ours-arm's: one step = the whole per-GPU batch (16 utterances take ~4 s on 16 cores); the
    # pipeline workloads (6 + 2 network forwards and two MVDRs per utterance) are bounded to 2 utterances per step
    utts = wl["B"] if wl["kind"] == "miso1" else 2
    cb = cpu_baseline(wl, steps=steps, warmup=max(1, min(args.warmup, 2)), utts=utts)
    return {
        "impl": "reference",
        "metric": "frames/sec MISO-BF-MISO fwd 6ch/257bin at 1/2/4/8 GPU; SI-SDR vs ref",
        "value": cb["value"], "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": max(1, min(args.warmup, 2)),
        "ms_per_step": cb["seconds_best"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded random spectrograms, seeded random weights)",
        "config": {"workload": wl["desc"], "layout": wl["layout"], "frames": wl["T"], "bins": wl["F"], "mics": wl["M"],
                   "per_gpu_batch": wl["B"], "utterances_per_step": utts,
                   "note": "reference CPU algorithm (oracle port of model.py/tester.py; /root/reference is absent on the GPU "
                           f"box), all host threads, each step = {utts} utterance(s) of the workload in one batch"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--workload", default="miso1_paper", choices=sorted(WORKLOADS) + ["train_paper"])
    ap.add_argument("--conv-mode", default="bf16x3", choices=sorted(CONV_MODES),
                    help="compute path of the conv stack (default: the parity-grade tensor-core mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (profiling runs only)")
    args = ap.parse_args()
    if args.workload == "train_paper":
        if args.impl == "reference":
            raise SystemExit("--workload train_paper: the CPU leg is the line's cpu_baseline (tools/train_step.py --cpu-baseline)")
        import runpy
        sys.argv = [sys.argv[0], "--layout", "PAPER", "--batch", str(args.batch or 8), "--steps", str(args.steps), "--warmup",
                    str(args.warmup), "--conv-mode", args.conv_mode] + ([] if args.no_cpu_baseline else ["--cpu-baseline"])
        runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "train_step.py"), run_name="__main__")
        return
    wl = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        wl["B"] = args.batch
        wl["desc"] += f" [batch overridden to {args.batch}: not a bench line]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = run_reference(args, wl, rank, world)
    elif args.impl == "torch_gpu":
        if wl["kind"] != "miso1":
            raise SystemExit("--impl torch_gpu covers the MISO1 forward workloads")
        line = run_torch_gpu(args, wl, rank, world, local)
    else:
        from misonet_b200 import distributed as D
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
        D.init_from_env("nccl")
        line = run_ours(args, wl, rank, world, local)
        if world > 1:
            torch.distributed.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
