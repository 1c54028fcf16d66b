"""Log-Mel error against float64 of every contraction route -- and of the unmodified reference on this GPU, with and
without PyTorch's default TF32 convolutions -- on mirror-stress signals (strong content at the mirror frequency N/2 - k of
a nearly silent band).  Source of profiles/r02_precision.md.  Run on the GPU box: python tests/gpu_precision_diag.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import reconvat_b200 as R
from oracle.frontend import FrontEndOracle
MEL_KW = dict(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000, trainable_mel=False, trainable_STFT=False, verbose=False)
n = 64 * 512 + 1
t = np.arange(n) / 16000.0
rng = np.random.default_rng(0)
sig = {"chirp": 0.7 * np.sin(2 * np.pi * (50.0 + 3800.0 * t / t[-1]) * t),
       "tone7k": 0.9 * np.sin(2 * np.pi * 7000.0 * t),
       "tone7k+low-60dB": 0.9 * np.sin(2 * np.pi * 7000.0 * t) + 0.9e-3 * np.sin(2 * np.pi * 500.0 * t),
       "tone500+hi-60dB": 0.9 * np.sin(2 * np.pi * 500.0 * t) + 0.9e-3 * np.sin(2 * np.pi * 7000.0 * t),
       "tone500+hi-80dB": 0.9 * np.sin(2 * np.pi * 500.0 * t) + 0.9e-4 * np.sin(2 * np.pi * 7000.0 * t)}
a16 = np.stack([np.clip(np.round(s * 32768.0), -32768, 32767).astype(np.int16) for s in sig.values()])
dev = torch.device("cuda:0")
orc = FrontEndOracle()
x = a16[:, :-1].astype(np.float64) / 32768.0
lr = np.log(orc.mel_power(x, np.float64) + 1e-5)
l32 = np.log(orc.mel_power(x.astype(np.float32), np.float32).astype(np.float64) + 1e-5)
rel = lambda a: (np.abs(a - lr) / np.maximum(np.abs(lr), 1)).reshape(len(sig), -1).max(1)
print("%-28s" % "route", "  ".join("%-16s" % k for k in sig))
print("%-28s" % "fp32 oracle (numpy)", "  ".join("%-16.2e" % v for v in rel(l32)))
# the reference's own modules on this GPU (eager fp32, tf32 off / on)
try:
    from oracle import reference_path
    rp = reference_path.ReferenceHotPath(dev)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            mel = rp.spectrogram(torch.from_numpy(x.astype(np.float32)).to(dev))
        print("%-28s" % ("reference on GPU tf32=%s" % tf32), "  ".join("%-16.2e" % v for v in rel(np.log(mel.cpu().numpy().astype(np.float64) + 1e-5))))
except Exception as e:
    print("reference unavailable", e)
for name, env in (("twice-folded f16 (default)", {}), ("once-folded f16", {"RVB_NO_FOLD2": "1"}), ("once-folded tf32", {"RVB_STFT_OPERAND": "tf32"}),
                  ("unfolded tf32", {"RVB_NO_FOLD": "1"})):
    os.environ.update(env)
    m = R.Spectrogram.MelSpectrogram(**MEL_KW).to(dev)
    mel = m(torch.from_numpy(a16).to(dev)[:, :-1]).cpu().numpy().astype(np.float64)
    for k in env:
        os.environ.pop(k)
    print("%-28s" % name, "  ".join("%-16.2e" % v for v in rel(np.log(mel + 1e-5))))
