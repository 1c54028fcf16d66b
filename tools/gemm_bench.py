"""Times the front-end kernels one by one (CUDA events, L2 flushed between launches by rotating buffers).
Usage: python tools/gemm_bench.py [B]   (RVB_GEMM_1CTA=1 selects the one-CTA contraction kernel)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import reconvat_b200 as R                                    # noqa: E402
from reconvat_b200 import _lib                               # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
mel = R.Spectrogram.MelSpectrogram(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
                                   verbose=False).to(dev)
audio = [(torch.rand(B, 327680, device=dev) * 2 - 1) for _ in range(4)]
names = ["rvb_fold_split_f16", "rvb_stft_gemm_folded_f16", "rvb_stft_mel_folded_f16", "rvb_mel_project", "rvb_normalise",
         "rvb_logmel_minmax", "rvb_logmel_transpose"]
for i in range(3):
    mel.normalised_log_mel(audio[i % 4])
torch.cuda.synchronize()
log = _lib.record_events(names)
N = 20
for i in range(N):
    mel.normalised_log_mel(audio[i % 4])
torch.cuda.synchronize()
_lib.record_events(None)
gemm = "rvb_stft_mel_folded_f16" if log["rvb_stft_mel_folded_f16"] else "rvb_stft_gemm_folded_f16"
for n in names:
    if not log[n]:
        continue
    ms = sorted(s.elapsed_time(e) for s, e in log[n])
    print("%-28s median %.1f us  min %.1f us" % (n, 1e3 * ms[len(ms) // 2], 1e3 * ms[0]))
flops = 3 * B * 640 * 2 * 1024 * 2048
ms = sorted(s.elapsed_time(e) for s, e in log[gemm])[N // 2]
print("GEMM issued %.0f TFLOP/s (%s kernel), B=%d" % (flops / ms / 1e9, "1-CTA" if os.environ.get("RVB_GEMM_1CTA") else "CTA-pair", B))
