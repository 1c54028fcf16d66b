"""Diagnostic (GPU): which variant of oracle/tc_accumulate.py reproduces the kind::f16 contraction bit for bit?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reconvat_b200 as R
from reconvat_b200 import _lib, synth
from oracle import tc_accumulate as TC

dev = torch.device("cuda:0")
stft = R.Spectrogram.STFT(n_fft=2048, hop_length=512, window='hann', freq_scale='no', center=True, pad_mode='reflect',
                          sr=16000, trainable=False, output_format='Complex', verbose=False).to(dev)
L = 5 * 512
a16 = torch.from_numpy(synth.music_int16(L, 5)[None].copy()).to(dev)
out = stft(a16).cpu().numpy()
fd = stft._device_tables()["fold"]
mode, n_frames, _ = stft._geometry(L)
planes = torch.empty((2, 2, n_frames, 1024), dtype=torch.float16, device=dev)
row_inv = torch.empty((n_frames,), dtype=torch.float32, device=dev)
_lib.call("rvb_fold_split_f16_pcm16", _lib.ptr(a16, torch.int16), L, 1.0 / 32768.0, 1, L, stft.pad_amount, mode,
          2048, 512, n_frames, planes[0].data_ptr(), planes[1].data_ptr(), row_inv.data_ptr(), None)
torch.cuda.synchronize()
pl = planes.cpu().numpy().astype(np.float64)
bh, bl = fd["basis_hi"].cpu().numpy().astype(np.float64), fd["basis_lo"].cpu().numpy().astype(np.float64)
ng, nb = fd["n_gemm_bins"], fd["n_bins_pad"]
scale = (row_inv.cpu().numpy() * np.float32(fd["scale_inv"]))[:, None]
got = out[0, :ng, :, 0].T
print("row_inv", row_inv.cpu().numpy(), "scale_inv", fd["scale_inv"], "one_pass env", os.environ.get("RVB_FOLD_ONE_PASS"))
sel = slice(0, 256)



terms = {"hh": (pl[0, 0], bh[:ng][sel]), "hl": (pl[0, 0], bl[:ng][sel]), "lh": (pl[1, 0], bh[:ng][sel])}
for pe in ("operands", "normalised"):
    for guard in (2, 3, 1):
        acc = TC.split_product(pl[0, 0], pl[1, 0], bh[:ng][sel], bl[:ng][sel], 16, order=("hl", "lh", "hh"),
                               corrections_first=True, guard_bits=guard, product_exponent=pe)
        re = acc.astype(np.float32) * scale
        bad = int((re != got[:, sel]).sum())
        ulp = np.abs(re.view(np.int32).astype(np.int64) - got[:, sel].copy().view(np.int32).astype(np.int64))
        print("f16 contraction  product_exponent=%-10s guard=%d: %5d of %d differ, max %d ulp" % (pe, guard, bad, re.size, ulp.max()))
        sys.stdout.flush()

from reconvat_b200 import linear
g = torch.Generator().manual_seed(11)
m, n, k = 37, 70, 224
x = torch.randn(m, k, generator=g) * torch.exp(2 * torch.randn(m, 1, generator=g))
w = torch.randn(n, k, generator=g)
xa, wb = linear._split(x.to(dev)), linear._split(w.to(dev))
o2 = torch.empty((m, n), dtype=torch.float32, device=dev)
_lib.call("rvb_gemm_nt_tf32x3", xa[0].data_ptr(), xa[1].data_ptr(), m, wb[0].data_ptr(), wb[1].data_ptr(), n, 224,
          o2.data_ptr(), n, 1, 0)
torch.cuda.synchronize()
f64 = lambda t: t.cpu().numpy().astype(np.float64)
g2 = o2.cpu().numpy()
for pe in ("operands", "normalised"):
    for guard in (2, 3, 1):
        want = TC.split_product(f64(xa[0]), f64(xa[1]), f64(wb[0]), f64(wb[1]), k_per_mma=8, order=("hh", "hl", "lh"),
                                guard_bits=guard, product_exponent=pe).astype(np.float32)
        ulp = np.abs(want.view(np.int32).astype(np.int64) - g2.view(np.int32).astype(np.int64))
        print("tf32 gemm        product_exponent=%-10s guard=%d: %5d of %d differ, max %d ulp" % (pe, guard, int((want != g2).sum()), want.size, ulp.max()))
