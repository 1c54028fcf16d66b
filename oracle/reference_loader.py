"""Import the UNMODIFIED reference hot-path modules.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The tree is looked for at ``$RECONVAT_REFERENCE``, then
``/root/reference`` (build container), then ``oracle/_ref/reference`` -- the byte-for-byte snapshot that
``__graft_entry__.build()`` makes with ``oracle/ref_snapshot.py`` so that the reference's own Python travels to
the GPU box (git-ignored build output, never committed).  Used by ``oracle/make_golden.py``, by the tests
(which skip when no tree is found) and by ``bench.py``'s reference legs.

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

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("RECONVAT_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref", "reference")):
        if cand and os.path.isdir(os.path.join(cand, "model")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def is_snapshot():
    """True when the tree in use is the shipped oracle/_ref copy (GPU box), not /root/reference itself."""
    return os.path.realpath(REFERENCE_ROOT).startswith(os.path.realpath(os.path.join(_HERE, "_ref")))


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


_MODEL_MODULES = ("constants", "utils", "VAT", "self_attention_VAT", "UNet_onset", "onset_frame_VAT", "Segmentation",
                  "decoding")
_patched = {}


def load_patched(attention=False, decoding=False, batchnorm=False):
    """The same UNMODIFIED reference modules, imported a second time behind the product's seams: the package
    ``reconvat_b200`` registered as ``nnAudio`` before the import and ``reconvat_b200.install()`` rebinding the VAT
    classes / Normalization afterwards -- what a user's ``sitecustomize`` does.  Returns a namespace
    like :func:`load_reference`, whose ``UNet`` / ``UNet_Onset`` / ``OnsetsAndFrames_VAT_full`` then run on librvb.so.
    The module objects are distinct from the unpatched ones, so both flavours can live in one process."""
    key = (bool(attention), bool(decoding), bool(batchnorm))
    if key in _patched:
        return _patched[key]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import reconvat_b200

    names = ["model", "nnAudio", "nnAudio.utils", "nnAudio.librosa_functions", "nnAudio.Spectrogram", "matplotlib",
             "matplotlib.pyplot"]
    saved = {k: sys.modules.get(k) for k in names}
    saved_model = {k: v for k, v in sys.modules.items() if k.startswith("model.")}
    for k in saved_model:
        del sys.modules[k]
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
    sys.modules.pop("nnAudio", None)
    ns = types.SimpleNamespace()
    try:
        reconvat_b200.install_nnaudio()                           # seam 1: before `import model`
        for name in _MODEL_MODULES:
            setattr(ns, name, importlib.import_module("model." + name))
        ns.rebound = reconvat_b200.patch_reference(attention=attention, decoding=decoding, batchnorm=batchnorm)   # seam 2
        ns.Spectrogram = reconvat_b200.Spectrogram
    finally:
        ns._mods = {k: sys.modules.pop(k) for k in [k for k in sys.modules if k.startswith("model.")]}
        sys.modules.update(saved_model)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _patched[key] = ns
    return ns
