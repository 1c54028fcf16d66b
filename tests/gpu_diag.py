"""Verbose per-kernel diagnostics for a GPU box (writes human-readable text; not a test).
Usage: python tests/gpu_diag.py [stage ...]   stages: vat pad mel gemm frontend timing"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import reconvat_b200 as R                                    # noqa: E402
from reconvat_b200 import _lib, basis, synth                 # noqa: E402
from oracle.frontend import FrontEndOracle                   # noqa: E402
from oracle import vat as OV                                 # noqa: E402

dev = torch.device("cuda:0")


def relerr(a, b):
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1)).max())


def stage_vat():
    torch.manual_seed(0)
    x = torch.rand(4, 1, 640, 229)
    d = torch.randn_like(x)
    g = torch.randn_like(x) * 1e-7
    xa = torch.empty_like(x, device=dev)
    xd0, dd0 = x.to(dev), d.to(dev)
    _lib.call("rvb_vat_perturb", xd0.data_ptr(), dd0.data_ptr(), xa.data_ptr(), 4 * 640, 229, 1e-6, 1)
    ref = OV.perturb(x, d, 1e-6)
    print("perturb max abs err", float((xa.cpu() - ref).abs().max()))
    xd, dd, gd = x.to(dev), d.to(dev), g.to(dev)
    r = torch.empty_like(xd); xa2 = torch.empty_like(xd); dh = torch.empty_like(xd)
    flag = torch.zeros((), dtype=torch.int32, device=dev)
    _lib.call("rvb_vat_finalize", gd.data_ptr(), dd.data_ptr(), xd.data_ptr(), r.data_ptr(), xa2.data_ptr(),
              dh.data_ptr(), 4 * 640, 229, 1e-6, 2.0, 1e10, 1, flag.data_ptr())
    dp = OV.power_grad_closed_form(x, d, g, 1e-6)
    r_ref, xa_ref, dh_ref = OV.finalize(x, dp, 2.0)
    print("finalize r_adv row err", float(((r.cpu() - r_ref).norm(dim=-1) / 2).max()), "flag", flag.item())
    print("finalize x_adv abs err", float((xa2.cpu() - xa_ref).abs().max()), "dhat", float((dh.cpu() - dh_ref).abs().max()))
    p = torch.sigmoid(torch.randn(4, 640, 88) * 3); y = torch.sigmoid(torch.randn(4, 640, 88) * 3)
    gr = torch.empty_like(p, device=dev)
    pd, yd = p.to(dev), y.to(dev)
    _lib.call("rvb_bce_grad", pd.data_ptr(), yd.data_ptr(), gr.data_ptr(), p.numel(), None, 1.0)
    ref = OV.bce_mean_grad(p, y)
    print("bce_grad rel-to-max err", float((gr.cpu() - ref).abs().max() / ref.abs().max()))
    ws = torch.zeros(_lib.BCE_WORKSPACE_FLOATS, device=dev); loss = torch.zeros((), device=dev)
    _lib.call("rvb_bce_mean", pd.data_ptr(), yd.data_ptr(), p.numel(), loss.data_ptr(), ws.data_ptr())
    print("bce_mean", loss.item(), OV.bce_mean(p, y).item())


def stage_pad():
    a = torch.from_numpy(synth.to_float(np.stack([synth.white_int16(16385, 1), synth.white_int16(16385, 2)])))
    x = a[:, :-1]
    L = x.shape[1]
    rows = -(-(L + 2048) // 512)
    sig = torch.zeros((2, 2 * rows, 512), device=dev)
    xd = a.to(dev)[:, :-1]
    _lib.call("rvb_pad_split", xd.data_ptr(), xd.stride(0), 2, L, 1024, 0, sig[0].data_ptr(), sig[1].data_ptr(), rows, 512)
    p = np.pad(x.numpy(), [(0, 0), (1024, 1024)], mode="reflect")
    pp = np.zeros((2, rows * 512), np.float32); pp[:, :p.shape[1]] = p
    hi, lo = basis.tf32_split(pp)
    print("pad_split hi exact", np.array_equal(sig[0].cpu().numpy().reshape(2, -1), hi),
          "lo exact", np.array_equal(sig[1].cpu().numpy().reshape(2, -1), lo))


def stage_mel():
    fo = FrontEndOracle()
    rng = np.random.default_rng(0)
    P = (rng.standard_normal((3, 70, 1024)).astype(np.float32) ** 2) * 50          # time-major [b][t][k]
    lo, ln, w, k_end = basis.band_rows(fo.mel_basis)
    t = lambda v: torch.from_numpy(v).to(dev)
    Pd, lod, lnd, wd = t(P), t(lo), t(ln), t(w)
    ref = np.einsum("mk,btk->bmt", fo.mel_basis[:, :1024].astype(np.float64), P.astype(np.float64))
    out = torch.empty((3, 229, 70), device=dev)
    _lib.call("rvb_mel_project", Pd.data_ptr(), 3, 70, 1024, lod.data_ptr(), lnd.data_ptr(), wd.data_ptr(), w.shape[0],
              229, -1.0, 0, out.data_ptr(), None)
    print("mel bins-major rel err", float((np.abs(out.cpu().numpy() - ref) / np.abs(ref).max()).max()))
    out2 = torch.empty((3, 70, 229), device=dev)
    mm = torch.empty((3, 2), dtype=torch.int32, device=dev)
    _lib.call("rvb_mel_project", Pd.data_ptr(), 3, 70, 1024, lod.data_ptr(), lnd.data_ptr(), wd.data_ptr(), w.shape[0],
              229, 1e-5, 1, out2.data_ptr(), mm.data_ptr())
    lref = np.log(ref + 1e-5).transpose(0, 2, 1)
    print("logmel time-major err", relerr(out2.cpu().numpy(), lref))
    o2 = out2.cpu().numpy()
    _lib.call("rvb_normalise", out2.data_ptr(), out2.data_ptr(), 3, 70 * 229, mm.data_ptr())
    mn = o2.reshape(3, -1).min(1)[:, None, None]; mx = o2.reshape(3, -1).max(1)[:, None, None]
    nref = (o2 - mn) / (mx - mn)
    print("normalise exact", np.array_equal(out2.cpu().numpy(), nref), float(np.abs(out2.cpu().numpy() - nref).max()))


def stage_gemm():
    fo = FrontEndOracle()
    a = synth.to_float(np.stack([synth.white_int16(16385, 1), synth.music_int16(16385, 2)]))
    st = R.Spectrogram.STFT(n_fft=2048, hop_length=512, sr=16000, verbose=False).to(dev)
    t0 = time.time()
    c = st(torch.from_numpy(a).to(dev)[:, :-1], output_format="Complex")
    torch.cuda.synchronize()
    print("stft complex ran in %.3f s" % (time.time() - t0), tuple(c.shape))
    c = c.cpu().numpy()
    re64, im64 = fo.stft(a[:, :-1].astype(np.float64), np.float64)
    scale = np.abs(re64).max()
    ere = np.abs(c[..., 0] - re64); eim = np.abs(c[..., 1] + im64)
    print("re max abs err / scale", ere.max() / scale, "im", eim.max() / scale, "scale", scale)
    # per 128-bin tile and per 32-bin chunk error map
    for j in range(8):
        sl = slice(128 * j, 128 * j + 128)
        print("  tile %d: re err %.3e im err %.3e" % (j, ere[:, sl].max() / scale, eim[:, sl].max() / scale))
    print("  nyquist bin: re err %.3e im err %.3e" % (ere[:, 1024].max() / scale, eim[:, 1024].max() / scale))
    print("  per-frame re err (first 8 frames):", (ere.max(axis=(0, 1)) / scale)[:8])
    if ere.max() / scale > 1e-3:
        print("  sample got/ref re[0, :4, :4]:\n", c[0, :4, :4, 0], "\n", re64[0, :4, :4])


def stage_bias():
    """Signed error of the contraction against float64 truth (tensor-core accumulate truncation)."""
    fo = FrontEndOracle()
    a = synth.to_float(np.stack([synth.white_int16(16385, 1), synth.music_int16(16385, 2)]))
    re64, im64 = fo.stft(a[:, :-1].astype(np.float64), np.float64)
    for label, env in (("folded", None), ("direct", "1")):
        if env:
            os.environ["RVB_NO_FOLD"] = env
        st = R.Spectrogram.STFT(n_fft=2048, hop_length=512, sr=16000, verbose=False).to(dev)
        c = st(torch.from_numpy(a).to(dev)[:, :-1], output_format="Complex").cpu().numpy().astype(np.float64)
        os.environ.pop("RVB_NO_FOLD", None)
        big = np.abs(re64) > 0.1 * np.abs(re64).max()
        rel = (c[..., 0] - re64)[big] / re64[big]
        scale = np.abs(re64).max()
        print("%s: max|err|/scale re %.3e im %.3e ; signed rel err on large bins: mean %.3e  std %.3e"
              % (label, np.abs(c[..., 0] - re64).max() / scale, np.abs(c[..., 1] + im64).max() / scale, rel.mean(), rel.std()))


def stage_weak():
    """Where does the log-Mel error live?  Error of re/im by magnitude decade, folded vs direct, on a full
    music-like segment; and the worst log-Mel cell of each path."""
    fo = FrontEndOracle()
    a = synth.segments(2, "mixed", seed=3)[1:2]                  # the music-like one
    re64, im64 = fo.stft(a[:, :-1].astype(np.float64), np.float64)
    mag = np.sqrt(re64 ** 2 + im64 ** 2)[:, :1024]
    scale = mag.max()
    lm64 = fo.log_mel(a[:, :-1].astype(np.float64), np.float64)
    for label, env in (("folded", None), ("direct", "1")):
        if env:
            os.environ["RVB_NO_FOLD"] = env
        st = R.Spectrogram.STFT(n_fft=2048, hop_length=512, sr=16000, verbose=False).to(dev)
        mel = R.Spectrogram.MelSpectrogram(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
                                           verbose=False).to(dev)
        ad = torch.from_numpy(a).to(dev)
        c = st(ad[:, :-1], output_format="Complex").cpu().numpy().astype(np.float64)[:, :1024]
        lm = np.log(mel(ad[:, :-1]).cpu().numpy().astype(np.float64) + 1e-5)
        os.environ.pop("RVB_NO_FOLD", None)
        err = np.sqrt((c[..., 0] - re64[:, :1024]) ** 2 + (c[..., 1] + im64[:, :1024]) ** 2)
        print("-- %s" % label)
        for lo_, hi_ in ((1e-7, 1e-5), (1e-5, 1e-4), (1e-4, 1e-3), (1e-3, 1e-2), (1e-2, 1e-1), (1e-1, 1.01)):
            m = (mag >= lo_ * scale) & (mag < hi_ * scale)
            if m.any():
                print("   |X|/scale in [%.0e,%.0e): n=%8d  rms err/scale %.2e  max err/scale %.2e  max err/|X| %.2e"
                      % (lo_, hi_, m.sum(), np.sqrt((err[m] ** 2).mean()) / scale, err[m].max() / scale, (err[m] / mag[m]).max()))
        e = np.abs(lm - lm64) / np.maximum(np.abs(lm64), 1)
        b, mm, t = np.unravel_index(e.argmax(), e.shape)
        print("   worst log-Mel cell: band %d frame %d  err %.3e  value %.4f (truth %.4f)" % (mm, t, e.max(), lm[b, mm, t], lm64[b, mm, t]))
        print("   log-Mel err percentiles 50/99/99.9/max: %s" % np.percentile(e, [50, 99, 99.9, 100]))


def stage_frontend():
    fo = FrontEndOracle()
    a = synth.segments(2, "mixed", seed=3)
    mel = R.Spectrogram.MelSpectrogram(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
                                       verbose=False).to(dev)
    ad = torch.from_numpy(a).to(dev)
    mp = mel(ad[:, :-1]).cpu().numpy()
    lm = np.log(mp + 1e-5)
    lm64 = fo.log_mel(a[:, :-1].astype(np.float64), np.float64)
    lm32 = fo.log_mel(a[:, :-1])
    print("log-mel err vs f64 truth", relerr(lm, lm64), "oracle32 vs f64", relerr(lm32, lm64), "vs oracle32", relerr(lm, lm32))
    spec = mel.normalised_log_mel(ad).cpu().numpy()
    print("fused spec err vs oracle", float(np.abs(spec - fo.spec_for_model(a)).max()), spec.shape, spec.min(), spec.max())


def stage_timing():
    mel = R.Spectrogram.MelSpectrogram(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
                                       verbose=False).to(dev)
    for B in (8, 32):
        a = torch.rand(B, 327680, device=dev) * 2 - 1
        for _ in range(3):
            mel.normalised_log_mel(a)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            mel.normalised_log_mel(a)
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        print("front-end B=%d: %.3f ms/step -> %.0f audio-s/s; STFT issued %.1f TFLOP/s (3xTF32), algorithmic %.1f"
              % (B, ms, B * 20.48 / ms * 1e3, 3 * B * 640 * 2048 * 2048 * 2 / ms / 1e9, B * 640 * 2050 * 2048 * 2 / ms / 1e9))


if __name__ == "__main__":
    stages = sys.argv[1:] or ["vat", "pad", "mel", "gemm", "bias", "frontend", "timing"]
    for s in stages:
        print("==== %s ====" % s, flush=True)
        try:
            globals()["stage_" + s]()
        except Exception as e:                                # keep going: later stages may still be informative
            import traceback
            traceback.print_exc()
            print("STAGE %s FAILED: %s" % (s, e))
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("CUDA context is dead:", e2)
                break
        sys.stdout.flush()
