"""reconvat_b200 -- B200-native (sm_100a) Mel front-end + VAT perturbation loop of ReconVAT.

Public surface (mirrors the reference, see DESIGN.md / INTEGRATION.md):
    reconvat_b200.Spectrogram.STFT / MelSpectrogram      <- nnAudio.Spectrogram (model/Spectrogram.py)
    reconvat_b200.VAT.*                                  <- the stepwise_VAT / UNet_VAT families
    reconvat_b200.utils.Normalization                    <- model/utils.py:82-106
    reconvat_b200.install()                              <- rebinds the above inside the reference's modules
    reconvat_b200.pipeline.HotPathStep                   <- the whole step as CUDA graphs (capture / replay / run_host)
Caller-side rows of SURVEY.md 8f (opt-in):
    reconvat_b200.transcribe.whole_file_frontend         <- UNet.transcribe's front-end, sharded by time
    reconvat_b200.attention.MutliHeadAttention1D         <- the U-Net's local-window attention
    reconvat_b200.decoding.extract_notes_wo_velocity     <- model/decoding.py
The kernels live in csrc/ behind the C ABI of include/rvb.h; there is no CPU or eager fallback.
"""
from . import _lib, basis                                   # noqa: F401
from . import Spectrogram, VAT, utils                       # noqa: F401
from .install import install, install_nnaudio, patch_reference   # noqa: F401

__version__ = "0.1.0"
