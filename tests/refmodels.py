"""Helpers for the tests that run the reference's OWN networks (test-pyramid item iv).

``oracle.reference_loader`` imports the unmodified reference modules twice -- once as they are (nnAudio helpers
restated, the vendored model/Spectrogram.py as ``nnAudio.Spectrogram``) and once behind ``reconvat_b200.install()``.
Both flavours are built from the same seed, so their parameters are identical, and are driven with the same batches
and the same global-generator seed, so both draw the same ``d`` (model/self_attention_VAT.py:172).
"""
import contextlib

import numpy as np
import pytest
import torch

SEG = 327680
# (module, class, constructor kwargs) of BASELINE.json's configs 2 / 3 / 4
MODELS = {
    "unet": ("self_attention_VAT", "UNet", dict(ds_ksize=(2, 2), ds_stride=(2, 2), log=True, reconstruction=True,
                                                mode="imagewise", spec="Mel")),          # train_UNet_VAT.py:126
    "unet_onset": ("UNet_onset", "UNet_Onset", dict(ds_ksize=(2, 2), ds_stride=(2, 2), log=True, reconstruction=True,
                                                    mode="imagewise", spec="Mel")),      # train_UNet_Onset_VAT.py:110
    "onf": ("onset_frame_VAT", "OnsetsAndFrames_VAT_full", dict(input_features=229, output_features=88,
                                                                model_complexity=48, log=True, mode="imagewise",
                                                                spec="Mel")),  # train_baseline_onset_frame_VAT.py:109
}


def namespaces():
    from oracle import reference_loader as RL
    if not RL.available():
        pytest.skip("no reference tree: neither /root/reference nor the oracle/_ref snapshot made by "
                    "__graft_entry__.build() is present")
    return RL.load_reference(), RL.load_patched()


def build(ns, name, device, xi, eps, seed=0):
    modname, clsname, kw = MODELS[name]
    cls = getattr(getattr(ns, modname), clsname)
    torch.manual_seed(seed)
    model = cls(XI=xi, eps=eps, **kw)
    return model.to(device)


def batch(n, seed, device, frames=640):
    """(audio, onset, frame) as model/dataset.py hands them over: float32 audio = int16 / 32768, labels in {0, 1}."""
    from reconvat_b200 import synth
    length = frames * 512
    audio = np.stack([(synth.music_int16 if (b & 1) else synth.white_int16)(length, seed * 1000 + b) for b in range(n)])
    g = torch.Generator().manual_seed(seed)
    frame = (torch.rand(n, frames, 88, generator=g) > 0.95).float()
    onset = (torch.rand(n, frames, 88, generator=g) > 0.99).float()
    return {"audio": torch.from_numpy(synth.to_float(audio)).to(device), "onset": onset.to(device),
            "frame": frame.to(device)}


@contextlib.contextmanager
def deterministic():
    """fp32 everywhere (no TF32 in cuDNN / cuBLAS) and deterministic algorithms where PyTorch has them."""
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
             torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.use_deterministic_algorithms(True, warn_only=True)
    try:
        yield
    finally:
        torch.use_deterministic_algorithms(False)
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
         torch.backends.cudnn.benchmark) = saved


def row_err(a, b, eps):
    """max over rows of ||a - b||_2 / eps (the r_adv tolerance of SURVEY.md 8d)."""
    return float(((a - b).reshape(-1, a.shape[-1]).norm(dim=-1) / eps).max())
