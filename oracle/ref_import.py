"""Import the real reference (read-only, /root/reference) with the shims it needs.

TEST INFRASTRUCTURE ONLY, and BUILD-CONTAINER ONLY: ``/root/reference`` does not
exist on the GPU box, so nothing that runs there may call this.  Used by
``oracle/make_golden.py`` (fixture generation) and by ``-m "not gpu"`` tests that
skip themselves when the reference is absent.

Shims:
* ``soundfile`` / ``librosa`` / ``matplotlib`` are absent -> empty module stubs.
* ``np.complex`` was removed in numpy 1.24 (tester.py:1104) -> alias to ``complex``.
* numpy >= 2.0 changed ``solve`` broadcasting for stacked vector right-hand sides
  (tester.py:1222) -> wrap with an explicit trailing axis.
Nothing in the reference is modified or copied.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MISONET_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


_cache = {}


def load():
    """Returns a namespace with .model, .criterion, .tester (reference modules)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    import numpy as np
    for name in ("soundfile", "librosa", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(np, "complex"):
        np.complex = complex
    import importlib.util

    def _load(modname, alias):
        spec = importlib.util.spec_from_file_location(alias, os.path.join(REFERENCE_ROOT, modname + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    # the reference modules import each other by bare name (e.g. tester.py imports
    # nothing of model.py, but dataloader/data.py imports siblings); put the root on
    # sys.path only for the duration of the import.
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ns = types.SimpleNamespace()
        ns.model = _load("model", "_miso_reference_model")
        ns.criterion = _load("criterion", "_miso_reference_criterion")
        ns.tester = _load("tester", "_miso_reference_tester")
        ns.tester.solve = lambda a, b: np.linalg.solve(a, b[..., None])[..., 0]
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _cache["ns"] = ns
    return ns


def make_tester(model_sep=None, model_enh=None, num_spks=2, ref_ch=0):
    """Tester_Enhance without its constructor (which wants datasets and directories):
    only the attributes the hot-path methods read (tester.py:1014-1244)."""
    ns = load()
    t = ns.tester.Tester_Enhance.__new__(ns.tester.Tester_Enhance)
    t.model_sep = model_sep
    t.model = model_enh
    t.num_spks = num_spks
    t.ref_ch = ref_ch
    return t


def make_stft_host(nperseg=256, noverlap=192, fs=8000):
    """An object exposing the reference STFT method (tester.py:992-1012) and scale
    (tester.py:818-819 region; same formula as dataloader/data.py:37-38)."""
    import numpy as np
    import scipy.signal
    ns = load()
    t = ns.tester.Tester_Enhance.__new__(ns.tester.Tester_Enhance)
    t.fs, t.window, t.nperseg, t.noverlap = fs, "hann", nperseg, noverlap
    hann = scipy.signal.get_window("hann", nperseg)
    t.scale = np.sqrt(1.0 / hann.sum() ** 2)
    return t
