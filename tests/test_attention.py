"""Local-window attention of the U-Net: oracle restatement pinned to the unmodified reference's
outputs (CPU), and the CUDA kernels (forward + backward) against both."""
import numpy as np
import pytest
import torch

from oracle import attention as OA
from reconvat_b200.standin import _hash_normal

CASES = ["small", "nopos", "unet"]


def _inputs(g, tag):
    B, L, fin, cout, W, G, pos = [int(v) for v in g[tag + "_dims"]]
    hn = lambda n, seed, shape: torch.tensor(_hash_normal(n, seed).reshape(shape), dtype=torch.float32)
    w = [hn(cout * fin, s, (cout, fin)) / np.sqrt(fin) for s in (301, 302, 303)]
    rel = hn(cout * W, 304, (1, cout, W)) if pos else None
    x = hn(B * L * fin, 305, (B, L, fin))
    go = hn(B * L * cout, 306, (B, L, cout))
    return (B, L, fin, cout, W, G, pos), w, rel, x, go


@pytest.mark.parametrize("tag", CASES)
def test_oracle_attention_matches_reference_golden(golden, tag):
    g = golden["attention"]
    (B, L, fin, cout, W, G, pos), w, rel, x, go = _inputs(g, tag)
    x = x.requires_grad_(True)
    if rel is not None:
        rel = rel.requires_grad_(True)
    out, att = OA.local_attention(x, w[0], w[1], w[2], rel, G, W)
    assert out.shape == (B, L, cout) and att.shape == (B, L, G, W)
    assert np.abs(out.detach().numpy() - g[tag + "_out"]).max() < 1e-5 * np.abs(g[tag + "_out"]).max()
    assert np.abs(att.detach().numpy() - g[tag + "_att"]).max() < 1e-6
    out.backward(go)
    assert np.abs(x.grad.numpy() - g[tag + "_dx"]).max() < 1e-5 * np.abs(g[tag + "_dx"]).max()
    if pos:
        assert np.abs(rel.grad.numpy() - g[tag + "_drel"]).max() < 1e-5 * np.abs(g[tag + "_drel"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("proj", ["tc", "torch"])
@pytest.mark.parametrize("tag", CASES)
def test_attention_kernels_match_reference_golden(golden, tag, proj, monkeypatch):
    """reconvat_b200.attention.MutliHeadAttention1D (forward, and backward to x, the projections and rel) against the
    outputs the UNMODIFIED reference class produced for the same weights and input -- with the projections on the
    tensor cores (3xTF32: 22 operand bits, energies of +-50 pass their 1e-5 through exp) and as nn.Linear."""
    import reconvat_b200.attention as A
    monkeypatch.setenv("RVB_ATTN_PROJ", proj)
    out_tol = 5e-5 if proj == "tc" else 2e-5
    dev = torch.device("cuda:0")
    g = golden["attention"]
    (B, L, fin, cout, W, G, pos), w, rel, x, go = _inputs(g, tag)
    m = A.MutliHeadAttention1D(fin, cout, W, position=pos, groups=G).to(dev)
    with torch.no_grad():
        m.W_q.weight.copy_(w[0]); m.W_k.weight.copy_(w[1]); m.W_v.weight.copy_(w[2])
        if pos:
            m.rel.copy_(rel)
    xd = x.to(dev).requires_grad_(True)
    out, att = m(xd)
    rel_err = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert out.shape == (B, L, cout) and att.shape == (B, L, G, W) and not att.requires_grad
    assert rel_err(out.detach().cpu().numpy(), g[tag + "_out"]) < out_tol
    # energies reach +-50 at the U-Net's head size: their fp32 summation-order rounding (~1e-5) passes through exp
    assert np.abs(att.cpu().numpy() - g[tag + "_att"]).max() < 5e-5
    assert torch.allclose(att.sum(-1), torch.ones_like(att[..., 0]), atol=1e-5)
    out.backward(go.to(dev))
    assert rel_err(xd.grad.cpu().numpy(), g[tag + "_dx"]) < 5e-5
    if pos:
        assert rel_err(m.rel.grad.cpu().numpy(), g[tag + "_drel"]) < 5e-5
    if tag != "unet":
        for nm in ("W_q", "W_k", "W_v"):
            assert rel_err(getattr(m, nm).weight.grad.cpu().numpy(), g[tag + "_d" + nm]) < 5e-5
    # state_dict interchange with the reference class: same parameter names and shapes
    assert sorted(m.state_dict()) == sorted(["W_k.weight", "W_q.weight", "W_v.weight"] + (["rel"] if pos else []))


@pytest.mark.gpu
def test_attention_full_size_against_oracle_and_memory():
    """Spec2Roll.lstm1 at full size (B=2, L=640, 229 -> 916, W=31, 4 heads): forward/backward against the oracle run on
    the CPU, and the point of the kernel -- no (B, L, C, W) tensor: peak memory stays a small multiple of q/k/v."""
    import reconvat_b200.attention as A
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = A.MutliHeadAttention1D(229, 916, 31, position=True, groups=4)
    x = torch.rand(2, 640, 229)
    go = torch.randn(2, 640, 916)
    xr = x.clone().requires_grad_(True)
    rel = m.rel.detach().clone().requires_grad_(True)
    o_ref, a_ref = OA.local_attention(xr, m.W_q.weight.detach(), m.W_k.weight.detach(), m.W_v.weight.detach(), rel, 4, 31)
    o_ref.backward(go)
    m = m.to(dev)
    xd = x.to(dev).requires_grad_(True)
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    out, att = m(xd)
    out.backward(go.to(dev))
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    rel_err = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel_err(out.detach().cpu(), o_ref.detach()) < 2e-5
    assert float((att.cpu() - a_ref.detach()).abs().max()) < 5e-5
    assert rel_err(xd.grad.cpu(), xr.grad) < 5e-5
    assert rel_err(m.rel.grad.cpu(), rel.grad) < 5e-5
    unfolded = 2 * 640 * 916 * 31 * 4                        # ONE of the reference's unfolded k / v tensors: 145 MB
    assert peak < 0.5 * unfolded, (peak, unfolded)


@pytest.mark.gpu
def test_attention_rejects_what_it_does_not_implement():
    import reconvat_b200.attention as A
    from reconvat_b200 import _lib
    with pytest.raises(NotImplementedError):
        A.MutliHeadAttention1D(8, 16, 3, stride=2)
    with pytest.raises(NotImplementedError):
        A.MutliHeadAttention1D(8, 16, 3, bias=True)
    with pytest.raises(AssertionError):
        A.MutliHeadAttention1D(8, 16, 4)                      # even window, like the reference
    with pytest.raises(_lib.RvbError):
        A.MutliHeadAttention1D(8, 16, 3)(torch.zeros(1, 5, 8))   # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("m,k,n", [(640, 229, 916), (300, 64, 40), (20480, 229, 916), (129, 33, 257)])
def test_tensor_core_projections_match_float64(m, k, n):
    """reconvat_b200.linear.projections (3xTF32 tcgen05 GEMM + tf32 split / transpose kernels): forward and the
    autograd backward (dx, dW) against float64 products; the error of a 3xTF32 contraction is ~2^-21 of the operand
    scale, like fp32 SGEMM's own accumulation rounding.  Shapes: one segment, ragged everything, the B = 32 layer, and
    sizes one past the tile edges."""
    from reconvat_b200 import linear
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(m + k + n)
    x = torch.randn(m, k, generator=g)
    ws = [torch.randn(n, k, generator=g) / np.sqrt(k), torch.randn(n // 2 + 1, k, generator=g) / np.sqrt(k)]
    gos = [torch.randn(m, w.shape[0], generator=g) for w in ws]
    xd = x.to(dev).requires_grad_(True)
    wd = [w.to(dev).requires_grad_(True) for w in ws]
    ys = linear.projections(xd, wd)
    torch.autograd.backward(ys, [go.to(dev) for go in gos])
    x64 = x.double().requires_grad_(True)
    w64 = [w.double().requires_grad_(True) for w in ws]
    y64 = [x64 @ w.t() for w in w64]
    torch.autograd.backward(y64, [go.double() for go in gos])
    rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())
    for y, yr in zip(ys, y64):
        assert y.shape == yr.shape and rel(y.detach(), yr.detach()) < 5e-6
    assert rel(xd.grad, x64.grad) < 5e-6
    for a, b in zip(wd, w64):
        assert rel(a.grad, b.grad) < 5e-6
    # the planes: hi is a tf32 number, hi + lo reconstructs x to 2^-21, padding columns are zero
    hi, lo = linear._split(xd.detach())
    assert hi.shape == (m, (k + 31) // 32 * 32) and (hi.view(torch.int32) & 0x1FFF).abs().max() == 0
    assert float((hi[:, :k] + lo[:, :k] - xd.detach()).abs().max()) <= 2.0 ** -20 * float(x.abs().max())
    assert float(hi[:, k:].abs().max() if hi.shape[1] > k else 0) == 0
    ht, lt = linear._split(xd.detach(), transpose=True)
    assert ht.shape[0] == k and torch.equal(ht[:, :m], hi[:, :k].t()) and torch.equal(lt[:, :m], lo[:, :k].t())


@pytest.mark.gpu
def test_attention_projection_paths_agree(monkeypatch):
    """MutliHeadAttention1D with the tensor-core projections vs with nn.Linear: same outputs and gradients to fp32
    rounding; a 3-D (B, L, C) input and the zero rows of the padding project to zero either way."""
    import reconvat_b200.attention as A
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = A.MutliHeadAttention1D(229, 916, 31, position=True, groups=4).to(dev)
    x = torch.randn(2, 70, 229, device=dev)
    go = torch.randn(2, 70, 916, device=dev)
    res = {}
    for mode in ("tc", "torch"):
        monkeypatch.setenv("RVB_ATTN_PROJ", mode)
        m.zero_grad()
        xi = x.clone().requires_grad_(True)
        out, att = m(xi)
        out.backward(go)
        res[mode] = (out.detach(), att, xi.grad, m.W_q.weight.grad.clone(), m.W_v.weight.grad.clone(), m.rel.grad.clone())
    for a, b in zip(res["tc"], res["torch"]):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max())


@pytest.mark.gpu
def test_tf32x3_gemm_equals_the_tensor_core_model_bit_for_bit():
    """rvb_gemm_nt_tf32x3 (one accumulator per element: k_split = 1) on random split operands against the CPU model of
    the tensor core's accumulation (oracle/tc_accumulate.py; hh, hl, lh per block of 8 terms): every bit."""
    from oracle import tc_accumulate as TC
    from reconvat_b200 import _lib, linear
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    m, n, k = 37, 70, 224
    x = torch.randn(m, k, generator=g) * torch.exp(2 * torch.randn(m, 1, generator=g))
    w = torch.randn(n, k, generator=g)
    xa, wb = linear._split(x.to(dev)), linear._split(w.to(dev))
    out = torch.empty((m, n), dtype=torch.float32, device=dev)
    _lib.call("rvb_gemm_nt_tf32x3", xa[0].data_ptr(), xa[1].data_ptr(), m, wb[0].data_ptr(), wb[1].data_ptr(), n, 224,
              out.data_ptr(), n, 1, 0)
    torch.cuda.synchronize()
    f64 = lambda t: t.cpu().numpy().astype(np.float64)
    want = TC.split_product(f64(xa[0]), f64(xa[1]), f64(wb[0]), f64(wb[1]), k_per_mma=8, order=("hh", "hl", "lh"))
    assert np.array_equal(out.cpu().numpy(), want.astype(np.float32))
    exact = (f64(xa[0]) + f64(xa[1])) @ (f64(wb[0]) + f64(wb[1])).T - f64(xa[1]) @ f64(wb[1]).T
    assert not np.array_equal(out.cpu().numpy(), exact.astype(np.float32))       # round-to-nearest would differ
