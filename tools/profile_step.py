"""Runs a few hot-path steps for ncu (no timing here: numbers taken under a profiler are never bench values).
Usage: python tools/profile_step.py [warmup] [steps] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_audio                                  # noqa: E402
from reconvat_b200.pipeline import HotPathStep                # noqa: E402
from reconvat_b200.standin import InjectedTranscriber         # noqa: E402

warmup = int(sys.argv[1]) if len(sys.argv) > 1 else 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda:0")
audio = [torch.from_numpy(a).to(dev) for a in make_audio(2, B, 0, pcm16=True)]      # PCM16, as bench.py
step = HotPathStep(InjectedTranscriber(B, seed=7).to(dev), dev)
for i in range(warmup + steps):
    step(audio[i % 2])
torch.cuda.synchronize()
step.vat_loss.check()
print("profile_step done: %d warm-up + %d steps, B=%d" % (warmup, steps, B))
