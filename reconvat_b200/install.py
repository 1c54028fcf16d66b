"""Make the reference's scripts pick up the B200 path without editing them.

Every model file of the reference does ``from nnAudio import Spectrogram`` and resolves its VAT class
and ``Normalization`` from its own module globals when the model is constructed
(model/self_attention_VAT.py:10,1027-1044; model/UNet_onset.py:10,354-363;
model/onset_frame_VAT.py:9,609-616).  Two seams follow:

1. :func:`install_nnaudio` registers this package's ``Spectrogram`` module as ``nnAudio.Spectrogram``
   in ``sys.modules`` -- call it before ``import model``;
2. :func:`patch_reference` rebinds the VAT / Normalization names inside the already imported
   reference modules -- call it after ``import model`` and before constructing the network.

``install()`` does (1) immediately and (2) for whatever reference modules are already imported; it is
idempotent, so calling it once before and once after ``import model`` covers both orders.  A
``sitecustomize`` that calls it makes ``python train_UNet_VAT.py with VAT=True`` run unmodified
(INTEGRATION.md).
"""
import sys
import types

from . import Spectrogram, VAT, basis, utils

_VAT_BINDINGS = {
    "model.VAT": {"stepwise_VAT": VAT.stepwise_VAT_vatpy},
    "model.self_attention_VAT": {"stepwise_VAT": VAT.stepwise_VAT, "UNet_VAT": VAT.UNet_VAT,
                                 "onset_frame_VAT": VAT.onset_frame_VAT},
    "model.UNet_onset": {"UNet_VAT": VAT.UNet_VAT_onset},
    "model.onset_frame_VAT": {"stepwise_VAT": VAT.stepwise_VAT_onf,
                              "stepwise_VAT_frame_stack": VAT.stepwise_VAT_frame_stack},
    "model.Segmentation": {"Seg_VAT": VAT.Seg_VAT},
}


def install_nnaudio():
    pkg = sys.modules.get("nnAudio")
    if pkg is None or getattr(pkg, "__reconvat_b200__", False) is False:
        pkg = types.ModuleType("nnAudio")
        pkg.__path__ = []
        pkg.__reconvat_b200__ = True
        sys.modules["nnAudio"] = pkg
    pkg.Spectrogram = Spectrogram
    sys.modules["nnAudio.Spectrogram"] = Spectrogram
    # the vendored model/Spectrogram.py star-imports these two (model/Spectrogram.py:15-16)
    u = types.ModuleType("nnAudio.utils")
    u.broadcast_dim = basis.broadcast_dim

    def create_fourier_kernels(n_fft, win_length=None, freq_bins=None, fmin=50, fmax=6000, sr=44100,
                               freq_scale='linear', window='hann', verbose=True):
        ks, kc, b2f, bl, wm = basis.fourier_basis(n_fft, win_length, freq_bins, window, freq_scale, fmin, fmax, sr)
        return ks[:, None, :], kc[:, None, :], b2f, bl, wm
    u.create_fourier_kernels = create_fourier_kernels
    u.__all__ = ["broadcast_dim", "create_fourier_kernels"]
    lf = types.ModuleType("nnAudio.librosa_functions")
    lf.mel = basis.mel_filterbank
    lf.__all__ = ["mel"]
    pkg.utils, pkg.librosa_functions = u, lf
    sys.modules["nnAudio.utils"] = u
    sys.modules["nnAudio.librosa_functions"] = lf
    return pkg


_ATTENTION_MODULES = ("model.self_attention_VAT", "model.UNet_onset", "model.onset_frame_VAT", "model.self_attention")


def _rebind_everywhere(name, new, defining_module):
    """Replace ``name`` in every loaded module that holds the reference's own object (scripts copy the names with
    ``from model import *``, transcribe_files.py:4, so the defining module is not the only holder)."""
    src = sys.modules.get(defining_module)
    old = getattr(src, name, None) if src is not None else None
    done = []
    if old is None or old is new:
        return done
    for modname, mod in list(sys.modules.items()):
        if mod is not None and getattr(mod, name, None) is old:
            setattr(mod, name, new)
            done.append((modname, name))
    return done


def patch_reference(attention=False, decoding=False, training=False, batchnorm=False):
    """Rebind VAT classes, Normalization and the Spectrogram module in every imported reference module.
    ``attention=True`` also rebinds ``MutliHeadAttention1D`` (the U-Net's sequence model) to the
    fused local-window attention: same parameters and outputs, no (B, L, C, W) unfolded tensors.
    ``decoding=True`` rebinds ``extract_notes_wo_velocity`` / ``notes_to_frames`` (model/decoding.py) wherever the
    reference's functions are held; they then expect the posteriors on the GPU, which is where ``UNet.transcribe``
    leaves them.
    ``training=True`` rebinds ``train_VAT_model`` (model/helper_functions.py:570-615) to ``reconvat_b200.training``'s:
    same arguments and arithmetic, no per-iteration host synchronisation, gradients averaged over the ranks when
    ``torch.distributed`` is initialised.
    ``batchnorm=True`` rebinds the name ``nn`` inside the reference's model files to a view of ``torch.nn`` whose
    ``BatchNorm2d`` is ``reconvat_b200.batchnorm.BatchNorm2d`` (same parameters, buffers and semantics; cuDNN's
    one-block-per-channel kernels are half of the U-Net's training iteration on a B200): models constructed afterwards
    use it, existing ones are converted with ``reconvat_b200.batchnorm.convert(model)``.
    Returns the list of (module, name) pairs that were rebound."""
    done = []
    if batchnorm:
        import torch.nn as _nn
        from . import batchnorm as _bn
        for modname, mod in list(sys.modules.items()):
            if mod is not None and modname.startswith("model.") and getattr(mod, "nn", None) is _nn:
                mod.nn = _bn.nn_proxy
                done.append((modname, "nn.BatchNorm2d"))
    if training:
        from . import training as _tr
        done += _rebind_everywhere("train_VAT_model", _tr.train_VAT_model, "model.helper_functions")
    if decoding:
        from . import decoding as _dec
        done += _rebind_everywhere("extract_notes_wo_velocity", _dec.extract_notes_wo_velocity, "model.decoding")
        done += _rebind_everywhere("notes_to_frames", _dec.notes_to_frames, "model.decoding")
    if attention:
        from . import attention as _att
        for modname in _ATTENTION_MODULES:
            mod = sys.modules.get(modname)
            if mod is not None and hasattr(mod, "MutliHeadAttention1D"):
                mod.MutliHeadAttention1D = _att.MutliHeadAttention1D
                done.append((modname, "MutliHeadAttention1D"))
    for modname, names in _VAT_BINDINGS.items():
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for name, cls in names.items():
            setattr(mod, name, cls)
            done.append((modname, name))
    for modname, mod in list(sys.modules.items()):
        if mod is None or not (modname == "model" or modname.startswith("model.")):
            continue
        if getattr(mod, "Normalization", None) is not None and mod.Normalization is not utils.Normalization:
            mod.Normalization = utils.Normalization
            done.append((modname, "Normalization"))
        if hasattr(mod, "Spectrogram") and isinstance(getattr(mod, "Spectrogram"), types.ModuleType) \
                and modname != "model.Spectrogram":
            mod.Spectrogram = Spectrogram
            done.append((modname, "Spectrogram"))
    pkg = sys.modules.get("model")
    if pkg is not None and hasattr(pkg, "stepwise_VAT"):
        # package-level star-import order makes `model.stepwise_VAT` the self_attention_VAT one
        pkg.stepwise_VAT = VAT.stepwise_VAT
        pkg.UNet_VAT = VAT.UNet_VAT
        done.append(("model", "stepwise_VAT"))
    return done


def install(attention=False, decoding=False, training=False, batchnorm=False):
    install_nnaudio()
    return patch_reference(attention=attention, decoding=decoding, training=training, batchnorm=batchnorm)
