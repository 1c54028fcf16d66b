"""Import the UNMODIFIED reference hot-path modules from /root/reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Build-container only: the GPU
box has no /root/reference, so nothing that runs there may call this; it is
used by ``oracle/make_golden.py`` and by CPU tests that skip when the tree is
absent.

``import model`` fails in this image because ``model/__init__.py:2-7`` pulls in
sacred / mir_eval / mido / soundfile / matplotlib.  The hot-path modules
themselves need only torch, numpy, scipy, PIL and nnAudio, so:

1. a bare ``model`` package whose ``__path__`` is the reference directory is
   pre-registered, which skips ``model/__init__.py``;
2. ``nnAudio.utils`` / ``nnAudio.librosa_functions`` are provided by
   ``oracle/nnaudio_restate.py``, and ``nnAudio.Spectrogram`` is the
   reference's own vendored ``model/Spectrogram.py``.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RECONVAT_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


_cached = None


def load_reference():
    """Returns a namespace: .Spectrogram, .utils, .constants, .self_attention_VAT,
    .UNet_onset, .onset_frame_VAT, .VAT, .Segmentation  (the reference's own module objects).
    model/Segmentation.py imports matplotlib.pyplot at module level without using it on the VAT path: an empty
    stand-in module is registered when matplotlib is not installed."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    from . import nnaudio_restate as R

    saved = {k: sys.modules.get(k) for k in
             ("model", "nnAudio", "nnAudio.utils", "nnAudio.librosa_functions", "nnAudio.Spectrogram",
              "matplotlib", "matplotlib.pyplot")}
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        mpl.__path__ = []
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot})

    pkg = types.ModuleType("model")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "model")]
    sys.modules["model"] = pkg

    nna = types.ModuleType("nnAudio")
    nna.__path__ = []
    u = types.ModuleType("nnAudio.utils")
    u.broadcast_dim = R.broadcast_dim
    u.create_fourier_kernels = R.create_fourier_kernels
    u.__all__ = ["broadcast_dim", "create_fourier_kernels"]
    lf = types.ModuleType("nnAudio.librosa_functions")
    lf.mel = R.mel
    lf.__all__ = ["mel"]
    nna.utils, nna.librosa_functions = u, lf
    sys.modules.update({"nnAudio": nna, "nnAudio.utils": u, "nnAudio.librosa_functions": lf})

    ns = types.SimpleNamespace()
    try:
        ns.Spectrogram = importlib.import_module("model.Spectrogram")
        nna.Spectrogram = ns.Spectrogram
        sys.modules["nnAudio.Spectrogram"] = ns.Spectrogram
        for name in ("constants", "utils", "VAT", "self_attention_VAT", "UNet_onset", "onset_frame_VAT", "Segmentation",
                     "decoding"):
            setattr(ns, name, importlib.import_module("model." + name))
    finally:
        # leave sys.modules as we found it: the product installs its own
        # `nnAudio` / patches `model.*`, and tests exercise that separately.
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith("model.")]:
            if saved.get("model") is None:
                ns.__dict__.setdefault("_mods", {})[k] = sys.modules.pop(k)
    _cached = ns
    return ns
