"""CPU tests of the host-side logic: basis builders (bit-exact against the reference-generated golden
tables), banded filterbank, tf32 split, module surface / state dict, geometry and error behaviour,
the install() seam, the no-CPU-fallback rule, the injected transcriber."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import reconvat_b200 as R
from reconvat_b200 import basis

MEL_KW = dict(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
              trainable_mel=False, trainable_STFT=False, verbose=False)


def test_basis_bit_exact_vs_reference_golden(golden):
    g = golden["basis_16k"]
    m = R.Spectrogram.MelSpectrogram(**MEL_KW)
    rows = g["rows"]
    assert np.array_equal(m.stft.wsin[rows, 0].numpy(), g["wsin_rows"])
    assert np.array_equal(m.stft.wcos[rows, 0].numpy(), g["wcos_rows"])
    assert np.array_equal(m.stft.window_mask.numpy().reshape(-1), g["window_mask"])
    assert np.isclose(m.stft.wcos.double().abs().sum().item(), g["wcos_abs64"], rtol=1e-12)
    assert np.isclose(m.stft.wsin.double().abs().sum().item(), g["wsin_abs64"], rtol=1e-12)
    mb = np.zeros(tuple(g["mel_shape"]), np.float32)
    mb[g["mel_nz_m"], g["mel_nz_k"]] = g["mel_nz_v"]
    assert np.array_equal(m.mel_basis.numpy(), mb)


def test_basis_independent_of_the_oracle_restatement():
    from oracle import nnaudio_restate as NR
    for kw in (dict(sr=22050, n_fft=2048, n_mels=128), dict(sr=16000, n_fft=1024, n_mels=128, htk=True),
               dict(sr=16000, n_fft=512, n_mels=40, fmin=20, fmax=7000)):
        assert np.array_equal(basis.mel_filterbank(**kw), NR.mel(**kw))
    for fs in ("no", "linear", "log"):
        a = basis.fourier_basis(512, freq_bins=100, freq_scale=fs, fmin=50, fmax=6000, sr=22050)
        b = NR.create_fourier_kernels(512, freq_bins=100, freq_scale=fs, fmin=50, fmax=6000, sr=22050, verbose=False)
        assert np.array_equal(a[0], b[0][:, 0]) and np.array_equal(a[1], b[1][:, 0]) and np.array_equal(a[4], b[4])


def test_banded_filterbank_roundtrip_and_rejection():
    mb = basis.mel_filterbank(16000, 2048, 229, 30, 8000)
    band0, w0, w1, kb, ke = basis.banded_filterbank(mb)
    assert (kb, ke) == (4, 1024) and np.all(np.diff(band0[kb:ke]) >= 0)
    dense = np.zeros_like(mb)
    for k in range(kb, ke):
        if w0[k]:
            dense[band0[k], k] = w0[k]
        if w1[k]:
            dense[band0[k] + 1, k] = w1[k]
    assert np.array_equal(dense, mb)
    assert int((mb != 0).sum()) == 2025 and int((mb != 0).sum(0).max()) == 2
    bad = mb.copy()
    bad[100, 50] = 1.0                                   # third, non-adjacent weight in a column
    with pytest.raises(ValueError, match="not a banded"):
        basis.banded_filterbank(bad)


def test_band_rows_roundtrip_and_rejection():
    mb = basis.mel_filterbank(16000, 2048, 229, 30, 8000)
    lo, ln, w, k_end = basis.band_rows(mb)
    assert k_end == 1024 and w.shape == (27, 229) and int(ln.max()) == 27 and int(lo.min()) == 4
    dense = np.zeros_like(mb)
    for m in range(229):
        dense[m, lo[m]:lo[m] + ln[m]] = w[:ln[m], m]
    assert np.array_equal(dense, mb)
    wide = mb.copy()
    wide[0, 900] = 1.0                                    # support of 897 bins: not banded
    with pytest.raises(ValueError, match="not a banded"):
        basis.band_rows(wide)
    for kw in (dict(sr=22050, n_fft=2048, n_mels=128), dict(sr=16000, n_fft=1024, n_mels=128, htk=True)):
        assert basis.band_rows(basis.mel_filterbank(**kw))[2].shape[0] <= 64


def test_mel_epilogue_table_roundtrip_and_rejection():
    mb = basis.mel_filterbank(16000, 2048, 229, 30, 8000)
    tab = basis.mel_epilogue_table(mb, 1024)
    band0 = tab[:, 2].view(np.int32)
    assert tab.shape == (1024, 4) and np.all(np.diff(band0) >= 0) and band0.max() == 227
    dense = np.zeros((229, 1024), np.float32)
    for k in range(1024):
        if tab[k, 0] != 0:
            dense[band0[k], k] += tab[k, 0]
        if tab[k, 1] != 0:
            dense[band0[k] + 1, k] += tab[k, 1]
    assert np.array_equal(dense, mb[:, :1024])
    three = mb.copy(); three[5, 600] = 1e-3                  # bin 600 would feed three bands
    assert basis.mel_epilogue_table(three, 1024) is None
    assert basis.mel_epilogue_table(mb, 896) is None         # weight on bins the contraction does not produce
    wide = basis.mel_filterbank(16000, 2048, 6, 30, 8000)    # bands hundreds of bins wide: > 2 tiles
    assert basis.mel_epilogue_table(wide, 1024) is None


def test_f16_fold_operand_is_block_scaled_and_reconstructs():
    ks, kc, _, _, win = basis.fourier_basis(2048, sr=16000)
    wcos, wsin = kc * win, ks * win
    f16 = basis.fold_operand(wcos, wsin, operand="f16")
    t32 = basis.fold_operand(wcos, wsin, operand="tf32")
    assert f16["basis_hi"].dtype == np.float16 and f16["basis_lo"].dtype == np.float16
    assert f16["basis_hi"].shape == t32["basis_hi"].shape == (2048, 1024)
    hi = f16["basis_hi"].astype(np.float64)
    assert np.isfinite(hi).all() and 2.0 ** 14 <= np.abs(hi).max() <= 2.0 ** 15
    assert np.log2(f16["scale_inv"]) == round(np.log2(f16["scale_inv"]))
    rec16 = (hi + f16["basis_lo"].astype(np.float64)) * f16["scale_inv"]
    rec32 = t32["basis_hi"].astype(np.float64) + t32["basis_lo"].astype(np.float64)
    # both splits carry 22 bits: they agree to 2^-21 of the element, or 2^-38 absolute where the fp16 lo is subnormal
    assert (np.abs(rec16 - rec32) <= 2.0 ** -21 * np.abs(rec32) + 2.0 ** -38).all()
    with pytest.raises(ValueError):
        basis.fold_operand(wcos, wsin, operand="bf16")


def test_tf32_split_reconstructs_to_2_pow_minus_21():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-20, 5, 100000))).astype(np.float32)
    hi, lo = basis.tf32_split(x)
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
    rel = np.abs((hi.astype(np.float64) + lo) - x) / np.abs(x)
    assert rel.max() < 2.0 ** -21
    # ties round away from zero, like cvt.rna.tf32.f32
    t = np.array([1.0 + 2.0 ** -11, -(1.0 + 2.0 ** -11)], np.float32)
    assert np.array_equal(basis.tf32_round(t), np.array([1.0 + 2.0 ** -10, -(1.0 + 2.0 ** -10)], np.float32))


def test_gemm_operand_layout():
    rng = np.random.default_rng(1)
    wc = rng.standard_normal((257, 512)).astype(np.float32)
    ws = rng.standard_normal((257, 512)).astype(np.float32)
    hi, lo, n_gemm, left = basis.gemm_operand(wc, ws)
    assert hi.shape == (512, 512) and n_gemm == 256 and left == [256]
    full = hi.astype(np.float64) + lo
    assert np.allclose(full[0:128], wc[0:128], rtol=1e-6) and np.allclose(full[128:256], ws[0:128], rtol=1e-6)
    assert np.allclose(full[256:384], wc[128:256], rtol=1e-6) and np.allclose(full[384:512], ws[128:256], rtol=1e-6)
    hi, lo, n_gemm, left = basis.gemm_operand(wc[:100], ws[:100])
    assert hi.shape == (256, 512) and n_gemm == 100 and left == [] and not hi[100:128].any() and not hi[228:].any()


def test_fold_operand_reproduces_the_unfolded_sums():
    ks, kc, _, _, wm = basis.fourier_basis(512, window="hamming")          # w[0] = 0.08: exercises the p0 term
    wcos, wsin = kc * wm[None], ks * wm[None]
    fd = basis.fold_operand(wcos, wsin)
    assert fd is not None and fd["n_bins_pad"] == 256 and fd["leftover"] == [256] and abs(fd["w0"] - 0.08) < 1e-6
    rng = np.random.default_rng(0)
    p = rng.standard_normal(512)
    e = np.empty(256); o = np.zeros(256)
    e[:-1] = p[1:256] + p[511:256:-1]; e[-1] = p[256]
    o[:-1] = p[1:256] - p[511:256:-1]
    B = fd["basis_hi"].astype(np.float64) + fd["basis_lo"]
    re, im = B[:256] @ e + fd["w0"] * p[0], B[256:] @ o
    assert np.abs(re - wcos[:256].astype(np.float64) @ p).max() < 1e-5
    assert np.abs(im - wsin[:256].astype(np.float64) @ p).max() < 1e-5
    assert abs(fd["left_cos"][0].astype(np.float64) @ e + fd["w0"] * p[0] - wcos[256].astype(np.float64) @ p) < 1e-5
    # not symmetric about n_fft/2 -> no fold
    ks, kc, _, _, wm = basis.fourier_basis(512, freq_bins=100, freq_scale="linear")
    assert basis.fold_operand(kc * wm[None], ks * wm[None]) is None
    assert basis.fold_operand(wcos[:, :500], wsin[:, :500]) is None        # n_fft not a multiple of 64


def test_module_surface_and_state_dict():
    m = R.Spectrogram.MelSpectrogram(**MEL_KW)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {
        "mel_basis": (229, 1025), "stft.wsin": (1025, 1, 2048), "stft.wcos": (1025, 1, 2048),
        "stft.window_mask": (1, 2048, 1)}
    assert m.stft.stride == 512 and m.stft.pad_amount == 1024 and m.power == 2.0
    s = R.Spectrogram.STFT(n_fft=1024, verbose=False)          # hop defaults to win_length // 4
    assert s.stride == 256 and s.wsin.shape == (513, 1, 1024) and s.output_format == "Complex"
    si = R.Spectrogram.STFT(n_fft=512, iSTFT=True, verbose=False)
    assert si.kernel_sin_inv.shape == (512, 1, 512, 1)
    for kw in (dict(trainable=True),):
        with pytest.raises(NotImplementedError):
            R.Spectrogram.STFT(verbose=False, **kw)
    with pytest.raises(NotImplementedError):
        R.Spectrogram.MelSpectrogram(trainable_mel=True, verbose=False)
    with pytest.raises(NotImplementedError):
        s.inverse(None)


def test_geometry_matches_conv1d_arithmetic():
    s = R.Spectrogram.STFT(n_fft=2048, hop_length=512, verbose=False)
    assert s._geometry(327679) == (0, 640, 644)
    assert s._geometry(1025)[1] == 3
    with pytest.raises(AssertionError, match="shorter than reflect padding"):
        s._geometry(1000)
    with pytest.raises(RuntimeError, match="Padding size should be less"):
        s._geometry(1024)
    nc = R.Spectrogram.STFT(n_fft=512, hop_length=256, center=False, verbose=False)
    assert nc._geometry(8192) == (2, 31, 32)
    with pytest.raises(RuntimeError, match="Kernel size"):
        nc._geometry(300)


def test_no_cpu_fallback_anywhere():
    m = R.Spectrogram.MelSpectrogram(**MEL_KW)
    with pytest.raises(R._lib.RvbError, match="no CPU path"):
        m(torch.zeros(1, 4096))
    with pytest.raises(R._lib.RvbError, match="no CPU path"):
        m.normalised_log_mel(torch.zeros(1, 4096))
    with pytest.raises(R._lib.RvbError, match="no CPU path"):
        R.utils.Normalization("imagewise").transform(torch.zeros(1, 4, 4))
    with pytest.raises(R._lib.RvbError, match="no CPU path"):
        R.VAT.UNet_VAT(1e-6, 2.0, 1, False)(None, torch.zeros(1, 1, 4, 229))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 1, 4096))
    from reconvat_b200 import attention, decoding
    with pytest.raises(R._lib.RvbError, match="no CPU path"):
        attention.MutliHeadAttention1D(8, 16, 3)(torch.zeros(1, 5, 8))
    with pytest.raises(R._lib.RvbError, match="no CPU path"):
        decoding.extract_notes_wo_velocity(torch.zeros(4, 88), torch.zeros(4, 88))
    with pytest.raises(R._lib.RvbError, match="no CPU path"):
        R.utils.Normalization("framewise").transform(torch.zeros(1, 4, 4))
    # the product never imports the oracle (nor do the tools)
    import os
    pkg = os.path.dirname(R.__file__)
    for top in (pkg, os.path.join(os.path.dirname(pkg), "tools"), os.path.join(os.path.dirname(pkg), "include")):
        for root, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(root, f)).read().replace("oracle/cpu_path", "").replace("oracle/vat.py", "")
                    assert "oracle" not in text, os.path.join(root, f)


def test_install_rebinds_the_reference_modules(monkeypatch):
    fake = {}
    for name in ("model", "model.VAT", "model.self_attention_VAT", "model.UNet_onset", "model.onset_frame_VAT"):
        mod = types.ModuleType(name)
        fake[name] = mod
        monkeypatch.setitem(sys.modules, name, mod)
    fake["model"].stepwise_VAT = object
    fake["model.self_attention_VAT"].Normalization = object
    fake["model.self_attention_VAT"].Spectrogram = types.ModuleType("nnAudio.Spectrogram")
    for k in ("nnAudio", "nnAudio.Spectrogram", "nnAudio.utils", "nnAudio.librosa_functions"):
        monkeypatch.delitem(sys.modules, k, raising=False)
    done = R.install()
    import nnAudio
    from nnAudio import Spectrogram as S
    assert S is R.Spectrogram and nnAudio.Spectrogram.MelSpectrogram is R.Spectrogram.MelSpectrogram
    assert fake["model.self_attention_VAT"].UNet_VAT is R.VAT.UNet_VAT
    assert fake["model.self_attention_VAT"].stepwise_VAT is R.VAT.stepwise_VAT
    assert fake["model.UNet_onset"].UNet_VAT is R.VAT.UNet_VAT_onset
    assert fake["model.onset_frame_VAT"].stepwise_VAT is R.VAT.stepwise_VAT_onf
    assert fake["model.VAT"].stepwise_VAT is R.VAT.stepwise_VAT_vatpy
    assert fake["model.self_attention_VAT"].Normalization is R.utils.Normalization
    assert fake["model.self_attention_VAT"].Spectrogram is R.Spectrogram
    assert fake["model"].stepwise_VAT is R.VAT.stepwise_VAT
    assert ("model.UNet_onset", "UNet_VAT") in done
    from nnAudio.utils import create_fourier_kernels
    from nnAudio.librosa_functions import mel
    assert create_fourier_kernels(512, freq_scale="no")[0].shape == (257, 1, 512) and mel(16000, 512).shape == (128, 257)


def test_install_on_the_real_reference_classes():
    """The seam on the reference's OWN modules (not fakes): imported behind install(), ``UNet`` / ``UNet_Onset`` /
    ``OnsetsAndFrames_VAT_full`` construct our MelSpectrogram / VAT / Normalization, their state_dicts are
    interchangeable with the unpatched models' (transcribe_files.py:71 loads strictly) and, built from the same seed,
    bit-identical -- buffers (wsin, wcos, window_mask, mel_basis) included."""
    import torch
    from oracle import reference_loader as RL
    if not RL.available():
        pytest.skip("no reference tree (/root/reference or the oracle/_ref snapshot made by build())")
    ref, pat = RL.load_reference(), RL.load_patched()
    assert ("model.self_attention_VAT", "UNet_VAT") in pat.rebound and ("model.UNet_onset", "UNet_VAT") in pat.rebound
    assert sys.modules.get("model") is None or not hasattr(sys.modules["model"], "__reconvat_test__")
    cases = [("self_attention_VAT", "UNet", ((2, 2), (2, 2)), dict(spec="Mel", XI=1e-6, eps=2), R.VAT.UNet_VAT),
             ("UNet_onset", "UNet_Onset", ((2, 2), (2, 2)), dict(spec="Mel"), R.VAT.UNet_VAT_onset),
             ("onset_frame_VAT", "OnsetsAndFrames_VAT_full", (229, 88), {}, R.VAT.stepwise_VAT_onf)]
    for mod, cls, args, kw, vat_cls in cases:
        torch.manual_seed(0)
        a = getattr(getattr(ref, mod), cls)(*args, **kw)
        torch.manual_seed(0)
        b = getattr(getattr(pat, mod), cls)(*args, **kw)
        assert type(b.spectrogram) is R.Spectrogram.MelSpectrogram and type(b.vat_loss) is vat_cls
        assert type(b.normalize) is R.utils.Normalization
        assert type(a.spectrogram).__module__ == "model.Spectrogram"          # the unpatched flavour stays unpatched
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa), cls
        b.load_state_dict(sa, strict=True)
        a.load_state_dict(sb, strict=True)
        assert b.vat_loss.XI == a.vat_loss.XI and b.vat_loss.epsilon == a.vat_loss.epsilon


def test_reference_snapshot_recipe(tmp_path):
    """oracle/ref_snapshot.py copies byte for byte and verifies by sha256 (what build() ships to the GPU box)."""
    from oracle import ref_snapshot, reference_loader as RL
    if not RL.available():
        pytest.skip("no reference tree")
    dest = str(tmp_path / "reference")
    manifest = ref_snapshot.materialise(source=RL.REFERENCE_ROOT, dest=dest, quiet=True)
    assert manifest and "model/self_attention_VAT.py" in manifest["files"] and ref_snapshot.verify(dest)
    with open(os.path.join(dest, "model", "VAT.py"), "a") as f:
        f.write("# tampered\n")
    assert not ref_snapshot.verify(dest)


def test_install_opt_in_attention_and_decoding(monkeypatch):
    """install(attention=True, decoding=True): the U-Net's sequence model and the note decoder are rebound too --
    the decoder in every module that copied the reference's function with `from model import *`."""
    import reconvat_b200.attention as A
    import reconvat_b200.decoding as D
    fake = {}
    for name in ("model", "model.self_attention_VAT", "model.UNet_onset", "model.decoding", "some_script"):
        fake[name] = types.ModuleType(name)
        monkeypatch.setitem(sys.modules, name, fake[name])
    ref_attention, ref_extract, ref_frames = object(), (lambda *a: None), (lambda *a: None)
    fake["model.self_attention_VAT"].MutliHeadAttention1D = ref_attention
    fake["model.UNet_onset"].MutliHeadAttention1D = ref_attention
    for m in ("model.decoding", "model", "some_script"):
        fake[m].extract_notes_wo_velocity = ref_extract
        fake[m].notes_to_frames = ref_frames
    for k in ("nnAudio", "nnAudio.Spectrogram", "nnAudio.utils", "nnAudio.librosa_functions"):
        monkeypatch.delitem(sys.modules, k, raising=False)
    assert R.install() is not None and fake["model.UNet_onset"].MutliHeadAttention1D is ref_attention   # opt-in only
    done = R.install(attention=True, decoding=True)
    assert fake["model.self_attention_VAT"].MutliHeadAttention1D is A.MutliHeadAttention1D
    assert fake["model.UNet_onset"].MutliHeadAttention1D is A.MutliHeadAttention1D
    for m in ("model.decoding", "model", "some_script"):
        assert fake[m].extract_notes_wo_velocity is D.extract_notes_wo_velocity
        assert fake[m].notes_to_frames is D.notes_to_frames
    assert ("some_script", "extract_notes_wo_velocity") in done
    # training=True: the step driver (model/helper_functions.py:570), also where a script star-imported it
    import inspect
    import reconvat_b200.training as T
    ref_train = (lambda *a, **k: None)
    fake["model.helper_functions"] = types.ModuleType("model.helper_functions")
    monkeypatch.setitem(sys.modules, "model.helper_functions", fake["model.helper_functions"])
    fake["model.helper_functions"].train_VAT_model = ref_train
    fake["some_script"].train_VAT_model = ref_train
    assert R.install() is not None and fake["some_script"].train_VAT_model is ref_train
    R.install(training=True)
    assert fake["model.helper_functions"].train_VAT_model is T.train_VAT_model
    assert fake["some_script"].train_VAT_model is T.train_VAT_model
    assert list(inspect.signature(T.train_VAT_model).parameters)[:11] == [
        "model", "iteration", "ep", "l_loader", "ul_loader", "optimizer", "scheduler", "clip_gradient_norm", "alpha",
        "VAT", "VAT_start"]


def test_vat_constructor_signatures_mirror_the_reference():
    V = R.VAT
    assert V.stepwise_VAT_vatpy(1e-6, 2, 1).epsilon == 2
    assert V.stepwise_VAT(1e-6, 2, 1, False, binwise=False).XI == 1e-6
    assert V.UNet_VAT(1e-6, 2, 1, False, reconstruction=True).reconstruction is True
    assert V.UNet_VAT_onset(1e-6, 2, 1, False)._dict_loss == ("frame", "onset")
    assert V.stepwise_VAT_onf(1e-6, 0.1, 1, False)._heads == (2,)
    assert V.onset_frame_VAT(1e-6, 2, 1)._n_returns == 2
    with pytest.raises(NotImplementedError):
        V.UNet_VAT(1e-6, 2, n_power=2, KL_Div=False)          # the reference fails for n_power > 1
    assert V.UNet_VAT(1e-6, 2, 1, True).KL_Div is True        # binary_kl_div flavour
    assert V.stepwise_VAT(1e-6, 2, 1, False, binwise=True).binwise is True
    assert V.Seg_VAT(1e-6, 2, 1, False, reconstruction=False)._heads == (None,)
    assert V.stepwise_VAT_frame_stack(1e-6, 2, 1, "all")._scale == 1e20
    for cls in (V.UNet_VAT_onset, V.Seg_VAT):                 # the reference raises NameError in these two
        with pytest.raises(NotImplementedError):
            cls(1e-6, 2, 1, True)


def test_injected_transcriber_drives_the_reference_op_sequence_on_cpu():
    from oracle.cpu_path import CpuHotPath
    from reconvat_b200.standin import InjectedTranscriber
    m = InjectedTranscriber(1, frames=6, seed=3)
    x = torch.rand(1, 1, 6, 229)
    torch.manual_seed(0)
    loss, r_adv, dhat = CpuHotPath().vat(m, x)
    assert torch.allclose(r_adv.norm(dim=-1), torch.full((1, 1, 6), 2.0), rtol=1e-5)
    assert torch.isfinite(loss)


def test_batchnorm_seam_and_convert_on_the_cpu():
    """install(batchnorm=True) rebinds the name `nn` inside the reference's model files to a view of torch.nn whose
    BatchNorm2d is ours (everything else falls through); convert() swaps classes in place and keeps the state_dict;
    the module refuses CPU tensors (no fallback)."""
    import sys
    import types
    import torch
    import torch.nn as nn
    import reconvat_b200
    from reconvat_b200 import _lib, batchnorm
    fake = types.ModuleType("model.fake_unet")
    fake.nn = nn
    sys.modules["model.fake_unet"] = fake
    try:
        done = reconvat_b200.patch_reference(batchnorm=True)
        assert ("model.fake_unet", "nn.BatchNorm2d") in done
        assert fake.nn.BatchNorm2d is batchnorm.BatchNorm2d and fake.nn.Conv2d is nn.Conv2d and fake.nn.Module is nn.Module
        bn = fake.nn.BatchNorm2d(8, momentum=0.1)
        assert isinstance(bn, nn.BatchNorm2d) and sorted(bn.state_dict()) == sorted(nn.BatchNorm2d(8).state_dict())
        assert reconvat_b200.patch_reference(batchnorm=True).count(("model.fake_unet", "nn.BatchNorm2d")) == 0   # idempotent
    finally:
        del sys.modules["model.fake_unet"]
    net = nn.Sequential(nn.Conv2d(1, 4, 3), nn.BatchNorm2d(4), nn.Sequential(nn.BatchNorm2d(4), nn.BatchNorm1d(4)))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    batchnorm.convert(net)
    assert type(net[1]) is batchnorm.BatchNorm2d and type(net[2][0]) is batchnorm.BatchNorm2d and type(net[2][1]) is nn.BatchNorm1d
    assert all(torch.equal(v, sd[k]) for k, v in net.state_dict().items())
    with pytest.raises(_lib.RvbError):
        net[1](torch.zeros(2, 4, 5, 5))
