"""Times the front-end entry points one by one: each is re-launched back to back between one CUDA-event pair, rotating
over recorded working sets (> L2), behind a spin kernel so that the events bracket kernels and not launch gaps.
Usage: python tools/gemm_bench.py [B] [reps]
Switches: RVB_NO_FOLD2=1 (once-folded contraction), RVB_FOLD2_N64=1 (four-chain twice-folded kernel), RVB_EXP=<bits> (ablation build, make ABLATION=1)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import reconvat_b200 as R                                    # noqa: E402
from reconvat_b200 import _lib                               # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
mel = R.Spectrogram.MelSpectrogram(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
                                   verbose=False).to(dev)
n_rot = 4
audio = [(torch.randint(-32768, 32767, (B, 327680), device=dev, dtype=torch.int16)) for _ in range(n_rot)]
for i in range(3):
    mel.normalised_log_mel(audio[i % n_rot])
torch.cuda.synchronize()
recorded, pools = [], []
for i in range(n_rot):
    pool = torch.cuda.MemPool()
    pools.append(pool)
    with torch.cuda.use_mem_pool(pool):
        log = []
        _lib.record_calls(log)
        mel.normalised_log_mel(audio[i])
        _lib.record_calls(None)
    recorded.append(log)
torch.cuda.synchronize()
for idx, (name, _) in enumerate(recorded[0]):
    calls = [recorded[i][idx][1] for i in range(n_rot)]
    for c in calls:
        _lib.raw_call(name, c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(reps * 40e-6 * 1.9e9))
    e0.record()
    for r in range(reps):
        _lib.raw_call(name, calls[r % n_rot])
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    extra = ""
    if "stft" in name:
        k_len = 512 if "folded2" in name else 1024
        extra = "  issued %.0f TFLOP/s" % (3 * B * 640 * 2 * 2048 * k_len / us / 1e6)
    print("%-28s %.1f us%s   (B=%d, RVB_EXP=%s)" % (name, us, extra, B, os.environ.get("RVB_EXP", "0")))
