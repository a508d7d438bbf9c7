"""ctypes binding of the C-ABI library (include/misonet_b200.h).

There is deliberately no fallback: if ``lib/libmisonet_b200.so`` is missing or a call
fails, an exception is raised.  The product path never routes through ``oracle/`` or
through stock PyTorch operators.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmisonet_b200.so")

_lib = None

# name -> (restype, argtypes); must list every symbol include/misonet_b200.h declares
SIGNATURES = {
    "miso_abi_version": (c_int, []),
    "miso_last_error": (c_char_p, []),
    "miso_check_device": (c_int, []),
    "miso_launch_count": (c_uint64, []),
    "miso_prof_enable": (c_int, [c_int]),
    "miso_prof_collect": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(ctypes.c_double),
                                  POINTER(c_uint64)]),
    "miso_prof_collect2": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(ctypes.c_double),
                                   POINTER(ctypes.c_double), POINTER(c_uint64)]),
    "miso_prof_dump": (c_int, [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(c_int), c_int]),
    "miso_stft_num_frames": (c_int, [c_int, c_int, c_int]),
    "miso_stft_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "miso_istft_num_samples": (c_int, [c_int, c_int, c_int]),
    "miso_istft_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "miso_istft_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                               c_void_p]),
    "miso_wave_to_int16": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "miso_net_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), c_int, c_int]),
    "miso_net_destroy": (c_int, [c_void_p]),
    "miso_net_num_params": (c_int, [c_void_p]),
    "miso_net_param_key": (c_char_p, [c_void_p, c_int]),
    "miso_net_param_numel": (c_int64, [c_void_p, c_int]),
    "miso_net_set_param": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "miso_net_set_params": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "miso_net_set_mode": (c_int, [c_void_p, c_int]),
    "miso_net_set_graph": (c_int, [c_void_p, c_int]),
    "miso_debug_tc_trace": (c_int, [c_void_p, c_int, c_int]),
    "miso_net_check_shape": (c_int, [c_void_p, c_int, c_int]),
    "miso_net_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "miso_net_input_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "miso_net_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "miso_net_grad_numel": (c_int64, [c_void_p]),
    "miso_net_grad_buckets": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), c_int]),
    "miso_net_wait_grad_bucket": (c_int, [c_void_p, c_int, c_void_p]),
    "miso_net_train_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "miso_net_forward_train": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "miso_net_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "miso_grad_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "miso_loss_enhance_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "miso_upit_bwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p]),
    "miso_net_tap": (c_int64, [c_void_p, c_char_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "miso_pack_miso1": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, POINTER(c_int), c_int, c_void_p]),
    "miso_pack_miso3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "miso_unpack_complex": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "miso_pair_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "miso_pair_fwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "miso_perm_gather": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_int,
                                 c_int, c_void_p]),
    "miso_loss_enhance_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "miso_mvdr_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "miso_mvdr_tsplit": (c_int, [c_int, c_int]),
    "miso_mvdr_partial_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "miso_mvdr_scm": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_int,
                              c_int, c_int, c_void_p]),
    "miso_mvdr_weights": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_size_t,
                                  c_void_p]),
    "miso_mvdr_apply": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_void_p]),
    "miso_mvdr_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int,
                              c_int, c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
}


class MisoError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MisoError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C misonet_b200/csrc`.  misonet_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.miso_abi_version() != 1:
        raise MisoError(f"ABI version mismatch: library {lib.miso_abi_version()}, binding 1")
    _lib = lib
    return lib


def check(rc, what="misonet_b200 call"):
    if rc is not None and rc < 0:
        msg = load().miso_last_error()
        raise MisoError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
    return rc


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def require_cuda(t, name):
    if not t.is_cuda:
        raise MisoError(f"{name} must be a CUDA tensor: misonet_b200 has no CPU path (got device {t.device})")


_device_checked = set()


def check_device(device):
    """Fail loudly on anything that is not a B200-class (sm_100) device."""
    import torch
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _device_checked:
        return
    with torch.cuda.device(idx):
        check(load().miso_check_device(), "miso_check_device")
    _device_checked.add(idx)


def launch_count():
    return int(load().miso_launch_count())
