"""One forward + backward of reconvat_b200.batchnorm.BatchNorm2d on the U-Net's largest tensor (8, 16, 640, 229), NCHW and
channels_last, and one pass of the strict front-end -- the launches `ncu --set full -k regex:"bn_|fold_pair"` captures for
profiles/r02_bn_strict_ncu_full.txt."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RVB_BN_NHWC"] = "1"
import reconvat_b200 as R  # noqa: E402
from reconvat_b200 import batchnorm, synth  # noqa: E402

dev = torch.device("cuda:0")
for fmt in (torch.contiguous_format, torch.channels_last):
    bn = batchnorm.BatchNorm2d(16).to(dev)
    x = torch.randn(8, 16, 640, 229, device=dev).contiguous(memory_format=fmt).requires_grad_(True)
    dy = torch.randn(8, 16, 640, 229, device=dev).contiguous(memory_format=fmt)
    for _ in range(2):
        bn(x).backward(dy)
a16 = torch.from_numpy(np.stack([synth.music_int16(synth.SEGMENT_SAMPLES, 900 + b) for b in range(32)])).to(dev)
mel = R.Spectrogram.MelSpectrogram(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000, verbose=False,
                                   precision="strict").to(dev)
for _ in range(2):
    mel.normalised_log_mel(a16)
torch.cuda.synchronize()
print("ok")
