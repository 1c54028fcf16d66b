"""How much do the contraction and the HBM kernels slow each other down when they run side by side?
Stream A re-launches the contraction back to back; stream B re-launches the HBM kernels of the step in a loop.
Prints the contraction's average launch time alone / beside B, and B's loop time alone / beside A.
Usage: python tools/contention_bench.py [B] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_audio                                  # noqa: E402
from reconvat_b200 import _lib                               # noqa: E402
from reconvat_b200.pipeline import HotPathStep                # noqa: E402
from reconvat_b200.standin import InjectedTranscriber         # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
n_rot = 4
audio = [torch.from_numpy(a).to(dev) for a in make_audio(n_rot, B, 0, pcm16=True)]
step = HotPathStep(InjectedTranscriber(B, seed=7).to(dev), dev)
for i in range(3):
    step(audio[i % n_rot])
torch.cuda.synchronize()
recorded, pools = [], []
for i in range(n_rot):
    pool = torch.cuda.MemPool()
    pools.append(pool)
    with torch.cuda.use_mem_pool(pool):
        log = []
        _lib.record_calls(log)
        step(audio[i])
        _lib.record_calls(None)
    recorded.append(log)
torch.cuda.synchronize()
step.vat_loss.check()
names = [n for n, _ in recorded[0]]
gi = next(i for i, n in enumerate(names) if "stft" in n)
hbm = [i for i, n in enumerate(names) if "stft" not in n]
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def run(gemm, others):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    with torch.cuda.stream(sa):
        torch.cuda._sleep(int(reps * 8 * 40e-6 * 1.9e9))
        ev[0].record()
        if gemm:
            for r in range(reps):
                _lib.raw_call(names[gi], recorded[r % n_rot][gi][1])
        ev[1].record()
    with torch.cuda.stream(sb):
        torch.cuda._sleep(int(reps * 8 * 40e-6 * 1.9e9))
        ev[2].record()
        if others:
            for r in range(reps):
                for i in hbm:
                    _lib.raw_call(names[i], recorded[r % n_rot][i][1])
        ev[3].record()
    torch.cuda.synchronize()
    return 1e3 * ev[0].elapsed_time(ev[1]) / reps, 1e3 * ev[2].elapsed_time(ev[3]) / reps


run(True, True)
g_alone, _ = run(True, False)
_, h_alone = run(False, True)
g_both, h_both = run(True, True)
print("contraction: %.1f us alone, %.1f us beside the HBM kernels" % (g_alone, g_both))
print("HBM kernels of one step (%s): %.1f us alone, %.1f us beside the contraction" % (", ".join(names[i] for i in hbm), h_alone, h_both))
print("serial %.1f us per step; side by side max(%.1f, %.1f) us" % (g_alone + h_alone, g_both, h_both))
