"""GPU parity of the VAT kernels / modules (through the C ABI) against the CPU oracle and the
reference-generated golden vectors.  Tolerances are the ones of SURVEY.md section 8d:
r_adv per-row ||delta||_2 / eps <= 1e-3, VAT loss rel <= 1e-3, BCE grad rel-to-max <= 1e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

R_ADV_TOL = 1e-3
LOSS_TOL = 1e-3
BCE_GRAD_TOL = 1e-5


@pytest.fixture(scope="module")
def R():
    import reconvat_b200
    return reconvat_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _spec_like(B, T, F=229, seed=0):
    from reconvat_b200 import synth
    x = synth.uniform01(B * T * F, seed).reshape(B, 1, T, F).astype(np.float32)
    x.reshape(B, -1)[:, 0] = 0.0
    x.reshape(B, -1)[:, 1] = 1.0
    return torch.from_numpy(x)


@pytest.mark.parametrize("B,T,F,clamp", [(2, 24, 229, 1), (4, 640, 229, 1), (1, 7, 229, 0), (3, 5, 88, 1), (2, 3, 400, 1)])
def test_perturb_matches_oracle(R, dev, B, T, F, clamp):
    from oracle import vat as OV
    x = _spec_like(B, T, F, 1)
    torch.manual_seed(3)
    d = torch.randn_like(x)
    out = torch.empty_like(x, device=dev)
    xd, dd = x.to(dev), d.to(dev)                          # keep the device copies alive across the launch
    R._lib.call("rvb_vat_perturb", xd.data_ptr(), dd.data_ptr(), out.data_ptr(), B * T, F, 1e-6, clamp)
    ref = OV.perturb(x, d, 1e-6, bool(clamp))
    # xi*dhat is ~1e-7: the sum rounds to x or x +- 1 ulp; allow one ulp of 1.0
    assert float((out.cpu() - ref).abs().max()) <= 1.2e-7
    if clamp:
        assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0


@pytest.mark.parametrize("B,T,F", [(2, 24, 229), (4, 640, 229), (2, 9, 88)])
@pytest.mark.parametrize("eps,scale,clamp", [(2.0, 1e10, 1), (1.3, 1e10, 1), (0.1, 1.0, 1), (2.0, 1.0, 0)])
def test_finalize_matches_oracle(R, dev, B, T, F, eps, scale, clamp):
    from oracle import vat as OV
    x = _spec_like(B, T, F, 2)
    torch.manual_seed(3)
    d = torch.randn_like(x)
    torch.manual_seed(4)
    g = torch.randn_like(x) * 1e-7
    xd, dd, gd = x.to(dev), d.to(dev), g.to(dev)
    r = torch.empty_like(xd); xa = torch.empty_like(xd); dh = torch.empty_like(xd)
    flag = torch.zeros((), dtype=torch.int32, device=dev)
    R._lib.call("rvb_vat_finalize", gd.data_ptr(), dd.data_ptr(), xd.data_ptr(), r.data_ptr(), xa.data_ptr(),
                dh.data_ptr(), B * T, F, 1e-6, eps, scale, clamp, flag.data_ptr())
    # reference arithmetic in float64 (the autograd chain of the reference, written out)
    dp = OV.power_grad_closed_form(x.double(), d.double(), g.double(), 1e-6, scale, bool(clamp))
    if clamp:   # the mask must be the fp32 one the forward pass saw
        s32 = x + 1e-6 * OV.l2_normalize(d)
        m = ((s32 >= 0) & (s32 <= 1)).double()
        n = d.double().norm(dim=-1, keepdim=True)
        gdv = 1e-6 * g.double() * m
        dp = (gdv / n - d.double() * ((gdv * d.double()).sum(-1, keepdim=True) / n ** 3)) * scale
    r_ref, xa_ref, dh_ref = OV.finalize(x.double(), dp, eps, bool(clamp))
    assert flag.item() == 0
    assert float(((r.cpu().double() - r_ref).norm(dim=-1) / eps).max()) < R_ADV_TOL * 1e-2
    assert float((dh.cpu().double() - dh_ref).norm(dim=-1).max()) < R_ADV_TOL * 1e-2
    assert float((xa.cpu().double() - xa_ref).abs().max()) < 1e-6
    # invariants the reference guarantees: ||r_adv||_row = eps, ||dhat||_row = 1
    assert torch.allclose(r.norm(dim=-1), torch.full_like(r[..., 0], eps), rtol=1e-5)
    assert torch.allclose(dh.norm(dim=-1), torch.ones_like(dh[..., 0]), rtol=1e-5)


def test_finalize_flags_nan_like_the_reference_assert(R, dev):
    x = _spec_like(1, 4, 229, 5).to(dev)
    d = torch.randn_like(x)
    g = torch.randn_like(x) * 1e-7
    g[0, 0, 2] = 0.0                       # a zero gradient row -> 0/0 in _l2_normalize -> NaN (self_attention_VAT.py:189)
    r = torch.empty_like(x); xa = torch.empty_like(x); dh = torch.empty_like(x)
    flag = torch.zeros((), dtype=torch.int32, device=dev)
    R._lib.call("rvb_vat_finalize", g.data_ptr(), d.data_ptr(), x.data_ptr(), r.data_ptr(), xa.data_ptr(),
                dh.data_ptr(), 4, 229, 1e-6, 2.0, 1e10, 1, flag.data_ptr())
    assert flag.item() & 1
    assert torch.isnan(r[0, 0, 2]).all() and not torch.isnan(r[0, 0, :2]).any()


@pytest.mark.parametrize("B,T,F", [(2, 24, 229), (32, 640, 229), (1, 3, 88), (3, 7, 400)])
@pytest.mark.parametrize("have_g", [True, False])
def test_finalize_stats_flavour(R, dev, B, T, F, have_g):
    """rvb_vat_finalize_stats: same r_adv / x_adv / d_hat bits as rvb_vat_finalize (rvb_vat_direct when g is NULL),
    the flag is WRITTEN by the last block (no zeroing, repeated calls on one workspace), and dhat_abs_mean is the
    reference's r_norm.abs().mean() (model/self_attention_VAT.py:1149) of the returned d_hat."""
    x = _spec_like(B, T, F, 6).to(dev)
    torch.manual_seed(7)
    d = torch.randn_like(x)
    g = torch.randn_like(x) * 1e-7
    n_rows = B * T
    outs = [[torch.empty_like(x) for _ in range(3)] for _ in range(2)]
    flag0 = torch.zeros((), dtype=torch.int32, device=dev)
    if have_g:
        R._lib.call("rvb_vat_finalize", g.data_ptr(), d.data_ptr(), x.data_ptr(), *[o.data_ptr() for o in outs[0]],
                    n_rows, F, 1e-6, 2.0, 1e10, 1, flag0.data_ptr())
    else:
        R._lib.call("rvb_vat_direct", d.data_ptr(), x.data_ptr(), *[o.data_ptr() for o in outs[0]], n_rows, F, 2.0, 1,
                    flag0.data_ptr())
    ws = torch.zeros((R._lib.vat_stats_workspace_bytes(n_rows) + 3) // 4, dtype=torch.int32, device=dev)
    mean = torch.full((), -1.0, device=dev)
    flag = torch.full((), 77, dtype=torch.int32, device=dev)          # garbage: the kernel must overwrite it
    vals = set()
    for rep in range(3):
        R._lib.call("rvb_vat_finalize_stats", g.data_ptr() if have_g else None, d.data_ptr(), x.data_ptr(),
                    *[o.data_ptr() for o in outs[1]], n_rows, F, 1e-6, 2.0, 1e10, 1, flag.data_ptr(), mean.data_ptr(),
                    ws.data_ptr(), ws.numel() * 4)
        assert flag.item() == 0 == flag0.item()
        vals.add(mean.item())
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert len(vals) == 1                                             # fixed-order sum: run-to-run identical
    want = outs[0][2].double().abs().mean().item()
    assert abs(vals.pop() - want) <= 2e-7 * want
    # NaN row: flag bit 0 is set, then cleared again by the next clean call on the same workspace
    if have_g:
        g2 = g.clone(); g2[0, 0, 1] = 0.0
        R._lib.call("rvb_vat_finalize_stats", g2.data_ptr(), d.data_ptr(), x.data_ptr(), *[o.data_ptr() for o in outs[1]],
                    n_rows, F, 1e-6, 2.0, 1e10, 1, flag.data_ptr(), mean.data_ptr(), ws.data_ptr(), ws.numel() * 4)
        assert flag.item() & 1 and torch.isnan(mean)
        R._lib.call("rvb_vat_finalize_stats", g.data_ptr(), d.data_ptr(), x.data_ptr(), *[o.data_ptr() for o in outs[1]],
                    n_rows, F, 1e-6, 2.0, 1e10, 1, flag.data_ptr(), mean.data_ptr(), ws.data_ptr(), ws.numel() * 4)
        assert flag.item() == 0
    # d_hat == NULL: nothing else changes (the mean still comes out of the same registers)
    r2, x2 = torch.empty_like(x), torch.empty_like(x)
    mean2 = torch.full((), -1.0, device=dev)
    R._lib.call("rvb_vat_finalize_stats", g.data_ptr() if have_g else None, d.data_ptr(), x.data_ptr(), r2.data_ptr(),
                x2.data_ptr(), None, n_rows, F, 1e-6, 2.0, 1e10, 1, flag.data_ptr(), mean2.data_ptr(), ws.data_ptr(),
                ws.numel() * 4)
    assert torch.equal(r2, outs[0][0]) and torch.equal(x2, outs[0][1]) and abs(mean2.item() - want) <= 2e-7 * want
    with pytest.raises(R._lib.RvbError, match="workspace"):
        R._lib.call("rvb_vat_finalize_stats", g.data_ptr(), d.data_ptr(), x.data_ptr(), *[o.data_ptr() for o in outs[1]],
                    n_rows, F, 1e-6, 2.0, 1e10, 1, flag.data_ptr(), mean.data_ptr(), ws.data_ptr(), 4)


def test_module_with_scratch_matches_module_without(R, dev):
    """vat.scratch = Scratch(dev) switches to the stats kernel and private workspaces: same r_adv, d_hat, loss; the
    deferred NaN assertion still fires."""
    from reconvat_b200.standin import StandInTranscriber
    gm = StandInTranscriber("unet", seed=2).to(dev)
    x = _spec_like(2, 16, 229, 8).unsqueeze(1).to(dev)
    res = []
    for use in (False, True):
        vat = R.VAT.UNet_VAT(0.1, 2.0, 1, False)
        if use:
            vat.scratch = R.VAT.Scratch(dev)
        torch.manual_seed(11)
        loss, r_adv, d_hat = vat(gm, x)
        res.append((loss.item(), r_adv.clone(), d_hat.clone(), vat.last_r_norm_mean))
        vat.check()
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    assert res[0][3] is None
    want = res[1][2].double().abs().mean().item()
    assert abs(res[1][3].item() - want) <= 2e-7 * want
    vat = R.VAT.UNet_VAT(0.1, 2.0, 1, False)
    vat.scratch = R.VAT.Scratch(dev, keep_d_hat=False)                # the direction itself is not stored
    torch.manual_seed(11)
    loss, r_adv, d_hat = vat(gm, x)
    assert d_hat is None and torch.equal(r_adv, res[1][1]) and loss.item() == res[1][0]
    assert vat.last_r_norm_mean.item() == res[1][3].item()


def test_bce_grad_and_mean(R, dev):
    import torch.nn.functional as F
    torch.manual_seed(5)
    p = torch.sigmoid(torch.randn(4, 640, 88) * 3)
    y = torch.sigmoid(torch.randn(4, 640, 88) * 3)
    p[0, 0, :4] = torch.tensor([0.0, 1.0, 1.0, 0.0])        # saturated posteriors: log clamp at -100, 1e-12 floor
    y[0, 0, :4] = torch.tensor([0.0, 1.0, 0.0, 1.0])
    pr = p.clone().requires_grad_(True)
    ref = F.binary_cross_entropy(pr, y)
    ref.backward()
    from reconvat_b200 import VAT
    pd = p.to(dev).requires_grad_(True)
    loss = VAT.bce_mean(pd, y.to(dev))
    loss.backward()
    assert abs(loss.item() - ref.item()) / ref.item() < 1e-6
    assert float((pd.grad.cpu() - pr.grad).abs().max() / pr.grad.abs().max()) < BCE_GRAD_TOL
    # upstream gradient is honoured, and the reduction is bit-reproducible
    pd2 = p.to(dev).requires_grad_(True)
    (VAT.bce_mean(pd2, y.to(dev)) * 3.0).backward()
    assert torch.allclose(pd2.grad, pd.grad * 3.0, rtol=1e-6)
    vals = {VAT.bce_mean(p.to(dev), y.to(dev)).item() for _ in range(5)}
    assert len(vals) == 1
    # odd sizes / unaligned views take the scalar path
    q = p.to(dev).reshape(-1)[1:1001]; z = y.to(dev).reshape(-1)[1:1001]
    assert abs(VAT.bce_mean(q, z).item() - F.binary_cross_entropy(q.cpu(), z.cpu()).item()) < 1e-6


FLAVOURS = {
    # tag: (class name, kwargs, stand-in convention, eps, d.grad scale, clamp)
    "unet": ("UNet_VAT", dict(epsilon=2, n_power=1, KL_Div=False), "unet", 2.0, 1e10, 1),
    "unet_eps13": ("UNet_VAT", dict(epsilon=1.3, n_power=1, KL_Div=False), "unet", 1.3, 1e10, 1),
    "stepwise_sa": ("stepwise_VAT", dict(epsilon=2, n_power=1, KL_Div=False), "stepwise", 2.0, 1.0, 1),
    "stepwise_vatpy": ("stepwise_VAT_vatpy", dict(epsilon=2, n_power=1), "stepwise", 2.0, 1.0, 0),
    "unet_onset": ("UNet_VAT_onset", dict(epsilon=2, n_power=1, KL_Div=False), "unet_onset", 2.0, 1e10, 1),
    "onf": ("stepwise_VAT_onf", dict(epsilon=0.1, n_power=1, KL_Div=False), "onf", 0.1, 1e10, 1),
    # flavours no shipped script selects
    "unet_kl": ("UNet_VAT", dict(epsilon=2, n_power=1, KL_Div=True), "unet", 2.0, 1e10, 1),
    "stepwise_sa_kl": ("stepwise_VAT", dict(epsilon=2, n_power=1, KL_Div=True), "stepwise", 2.0, 1.0, 1),
    "onf_kl": ("stepwise_VAT_onf", dict(epsilon=0.1, n_power=1, KL_Div=True), "onf", 0.1, 1e10, 1),
    "seg": ("Seg_VAT", dict(epsilon=2, n_power=1, KL_Div=False), "seg", 2.0, 1e10, 1),
    "stack_activation": ("stepwise_VAT_frame_stack", dict(epsilon=2, n_power=1, VAT_mode="activation"), "stack", 2.0, 1e20, 1),
    "stack_frame": ("stepwise_VAT_frame_stack", dict(epsilon=2, n_power=1, VAT_mode="frame"), "stack", 2.0, 1e20, 1),
    "stack_all": ("stepwise_VAT_frame_stack", dict(epsilon=2, n_power=1, VAT_mode="all"), "stack", 2.0, 1e20, 1),
}


@pytest.mark.parametrize("xi", [1e-6, 0.1])
@pytest.mark.parametrize("base", sorted(FLAVOURS))
def test_vat_kernels_match_reference_golden_with_injected_g(R, dev, golden, base, xi):
    """The reference's own (x, d, g) -> our perturb + finalize kernels -> the reference's r_adv / dhat.
    g = dL/dx_adv is injected because, with the shipped XI=1e-6, it is fp32-rounding-level sensitive to the
    device the network runs on ("Where g comes from")."""
    g = golden["vat_flavours"]
    tag = base if xi == 1e-6 else base + "_xi01"
    _, _, _, eps, scale, clamp = FLAVOURS[base]
    d = torch.from_numpy(g[tag + "_d"]).to(dev)
    x = torch.from_numpy(g["x"] if d.dim() == 4 else g["x"][:, 0]).to(dev).contiguous()
    gg = torch.from_numpy(g[tag + "_g"][0]).to(dev).contiguous()
    n_rows, F = x.numel() // 229, 229
    r = torch.empty_like(x); xa = torch.empty_like(x); dh = torch.empty_like(x)
    flag = torch.zeros((), dtype=torch.int32, device=dev)
    R._lib.call("rvb_vat_finalize", gg.data_ptr(), d.data_ptr(), x.data_ptr(), r.data_ptr(), xa.data_ptr(),
                dh.data_ptr(), n_rows, F, xi, eps, scale, clamp, flag.data_ptr())
    r_ref = torch.from_numpy(g[tag + "_r_adv"])
    assert flag.item() == 0
    assert float(((r.cpu() - r_ref).norm(dim=-1) / eps).max()) < R_ADV_TOL
    if tag + "_dhat" in g.files:
        assert float((dh.cpu() - torch.from_numpy(g[tag + "_dhat"])).norm(dim=-1).max()) < R_ADV_TOL
    # x_adv handed to the final network pass: clamp(x + r_adv, 0, 1) (no clamp in the model/VAT.py flavour)
    xa_ref = x.cpu() + r_ref
    if clamp:
        xa_ref = xa_ref.clamp(0, 1)
    assert float((xa.cpu() - xa_ref).abs().max()) < 1e-5


@pytest.mark.parametrize("base", sorted(FLAVOURS))
def test_vat_modules_match_reference_golden(R, dev, golden, base, monkeypatch):
    """Whole module forward (our kernels + the stand-in network on the GPU) against the outputs the
    UNMODIFIED reference produced for the same x, d and network, on the well-conditioned XI=0.1 twins."""
    from reconvat_b200.standin import StandInTranscriber
    g = golden["vat_flavours"]
    tag = base + "_xi01"
    cls, kw, conv, eps, _, _ = FLAVOURS[base]
    d = torch.from_numpy(g[tag + "_d"]).to(dev)
    x = torch.from_numpy(g["x"] if d.dim() == 4 else g["x"][:, 0]).to(dev)
    model = StandInTranscriber(conv, n_in=229, n_out=int(g["P"]), seed=3).to(dev)
    model.captured_grads = []
    vat = getattr(R.VAT, cls)(XI=0.1, strict=True, **kw)
    monkeypatch.setattr(torch, "randn_like", lambda t, **k: d.clone())     # inject the reference's d
    res = vat(model, x)
    vat_loss, r_adv = res[0], res[1]
    r_ref = torch.from_numpy(g[tag + "_r_adv"])
    gref = g[tag + "_g"][0]
    assert float(np.abs(model.captured_grads[0].cpu().numpy() - gref).max() / np.abs(gref).max()) < 1e-3
    assert float(((r_adv.cpu() - r_ref).norm(dim=-1) / eps).max()) < R_ADV_TOL
    if len(res) > 2:
        assert float((res[2].cpu() - torch.from_numpy(g[tag + "_dhat"])).norm(dim=-1).max()) < R_ADV_TOL
    else:
        assert tag + "_dhat" not in g.files
    if isinstance(vat_loss, dict):
        got = np.array([vat_loss["frame"].item(), vat_loss["onset"].item()])
        total = vat_loss["frame"] + vat_loss["onset"]
    else:
        got = np.array([vat_loss.item()])
        total = vat_loss
    assert np.allclose(got, g[tag + "_loss"], rtol=LOSS_TOL)
    # vat_loss carries a graph to the network parameters (it is summed into the training loss)
    model.captured_grads = None
    total.backward()
    wref = g[tag + "_wgrad"]
    assert float(np.abs(model.frame.weight.grad.cpu().numpy() - wref).max() / np.abs(wref).max()) < 1e-3
    assert not r_adv.requires_grad and r_adv.shape == x.shape


@pytest.mark.parametrize("xi", [1e-6, 0.1])
def test_binwise_kernels_reproduce_the_reference_bit_for_bit(R, dev, golden, xi):
    """binwise=True: the reference's d.grad is fp32 cancellation noise (oracle/vat.py:power_grad_binwise_sequence);
    the kernel executes the same IEEE op sequence, so with the reference's (x, d, g) the outputs are IDENTICAL."""
    g = golden["vat_flavours"]
    tag = "stepwise_sa_binwise" + ("" if xi == 1e-6 else "_xi01")
    x = torch.from_numpy(g["x"]).to(dev)
    d = torch.from_numpy(g[tag + "_d"]).to(dev)
    gg = torch.from_numpy(g[tag + "_g"][0]).to(dev).contiguous()
    r = torch.empty_like(x); xa = torch.empty_like(x); dh = torch.empty_like(x)
    flag = torch.zeros((), dtype=torch.int32, device=dev)
    R._lib.call("rvb_vat_finalize_binwise", gg.data_ptr(), d.data_ptr(), x.data_ptr(), r.data_ptr(), xa.data_ptr(),
                dh.data_ptr(), x.numel(), xi, 2.0, 1.0, 1, flag.data_ptr())
    assert flag.item() == 0
    assert torch.equal(r.cpu(), torch.from_numpy(g[tag + "_r_adv"]))
    assert torch.equal(dh.cpu(), torch.from_numpy(g[tag + "_dhat"]))
    x1 = torch.empty_like(x)
    R._lib.call("rvb_vat_perturb_binwise", x.data_ptr(), d.data_ptr(), x1.data_ptr(), x.numel(), xi, 1)
    want = (x.cpu() + xi * (d.cpu() / (d.cpu().abs() + 1e-8))).clamp(0, 1)
    assert torch.equal(x1.cpu(), want)


def test_binwise_module_and_n_power_zero(R, dev, golden):
    """Whole module with binwise=True on the GPU network.  g differs from the CPU run at rounding level and the
    direction amplifies that noise, so only what is well defined is compared: the loss, shapes, |d_hat| <= 1."""
    from reconvat_b200.standin import StandInTranscriber
    g = golden["vat_flavours"]
    x = torch.from_numpy(g["x"]).to(dev)
    model = StandInTranscriber("stepwise", n_in=229, n_out=int(g["P"]), seed=3).to(dev)
    vat = R.VAT.stepwise_VAT(0.1, 2, 1, False, binwise=True, strict=True)
    loss, r_adv, d_hat = vat(model, x)
    assert r_adv.shape == x.shape and float(d_hat.abs().max()) <= 1.0 and torch.allclose(r_adv, 2.0 * d_hat)
    assert np.isfinite(loss.item())
    loss.backward()
    assert model.frame.weight.grad is not None
    v0 = R.VAT.stepwise_VAT(0.1, 2, 0, False, binwise=True)
    torch.manual_seed(5)
    d = torch.randn_like(x)
    torch.manual_seed(5)
    _, r0, dh0 = v0(model, x)
    assert torch.equal(dh0, d / (d.abs() + 1e-8)) and torch.equal(r0, 2.0 * dh0)


def test_divergence_kernels_match_torch(R, dev):
    """rvb_div_mean / rvb_div_grad for the binary KL (batchmean) and MSE against the reference's torch expressions."""
    import torch.nn.functional as F
    from oracle import vat as OV
    from reconvat_b200 import VAT
    torch.manual_seed(6)
    p = torch.sigmoid(torch.randn(3, 40, 88) * 4)
    y = torch.sigmoid(torch.randn(3, 40, 88) * 4)
    p[0, 0, :6] = torch.tensor([0.0, 1.0, 1e-4, 0.9999, 5e-5, 0.99995])      # on and beyond the clamp edges
    for ours, ref in ((VAT.binary_kl_div, OV.binary_kl_div), (VAT.mse_mean, F.mse_loss)):
        pr = p.clone().requires_grad_(True)
        lr = ref(pr, y)
        lr.backward()
        pd = p.to(dev).requires_grad_(True)
        lo = ours(pd, y.to(dev))
        (lo * 2.0).backward()
        assert abs(lo.item() - lr.item()) <= 2e-6 * abs(lr.item())
        assert float((pd.grad.cpu() - 2.0 * pr.grad).abs().max() / pr.grad.abs().max()) < 1e-5


@pytest.mark.parametrize("xi", [0.1])
def test_vat_full_size_against_oracle(R, dev, xi):
    """B=4 x 640 x 229 (one BASELINE config-2 half batch): module vs oracle with the same d and network."""
    from oracle import vat as OV
    from reconvat_b200.standin import StandInTranscriber
    x = _spec_like(4, 640, 229, 9)
    model = StandInTranscriber("unet", seed=2)
    torch.manual_seed(11)
    d = torch.randn_like(x)
    l_ref, r_ref, dh_ref, g_ref = OV.vat_unet(lambda z: model.transcriber(z)[0], x, d, xi, 2.0)
    gm = StandInTranscriber("unet", seed=2).to(dev)
    vat = R.VAT.UNet_VAT(xi, 2.0, 1, False)
    orig = torch.randn_like
    try:
        torch.randn_like = lambda t, **k: d.to(dev)
        loss, r_adv, d_hat = vat(gm, x.to(dev))
    finally:
        torch.randn_like = orig
    vat.check()
    assert float(((r_adv.cpu() - r_ref).norm(dim=-1) / 2.0).max()) < R_ADV_TOL
    assert abs(loss.item() - l_ref.item()) / abs(l_ref.item()) < LOSS_TOL
    assert torch.allclose(r_adv.norm(dim=-1), torch.full_like(r_adv[..., 0], 2.0), rtol=1e-5)


def test_vat_shipped_xi_full_size_with_injected_g(R, dev):
    """Shipped hyper-parameters (XI=1e-6, eps=2) at full size: g from the CPU network injected into the
    device kernels, r_adv against the oracle."""
    from oracle import vat as OV
    from reconvat_b200.standin import StandInTranscriber
    x = _spec_like(4, 640, 229, 9)
    model = StandInTranscriber("unet", seed=2)
    torch.manual_seed(11)
    d = torch.randn_like(x)
    l_ref, r_ref, dh_ref, g_ref = OV.vat_unet(lambda z: model.transcriber(z)[0], x, d, 1e-6, 2.0)
    xd, dd, gd = x.to(dev), d.to(dev), g_ref.to(dev)
    r = torch.empty_like(xd); xa = torch.empty_like(xd); dh = torch.empty_like(xd)
    flag = torch.zeros((), dtype=torch.int32, device=dev)
    R._lib.call("rvb_vat_finalize", gd.data_ptr(), dd.data_ptr(), xd.data_ptr(), r.data_ptr(), xa.data_ptr(),
                dh.data_ptr(), 4 * 640, 229, 1e-6, 2.0, 1e10, 1, flag.data_ptr())
    assert float(((r.cpu() - r_ref).norm(dim=-1) / 2.0).max()) < R_ADV_TOL
    gm = StandInTranscriber("unet", seed=2).to(dev)
    with torch.no_grad():
        loss = R.VAT.bce_mean(gm.transcriber(xa)[0], gm.transcriber(xd)[0])
    assert abs(loss.item() - l_ref.item()) / abs(l_ref.item()) < LOSS_TOL


def test_vat_nan_assertion_synchronous_by_default_deferred_on_request(R, dev):
    from reconvat_b200.standin import StandInTranscriber

    class Dead(StandInTranscriber):           # a network whose output ignores x: g == 0 -> r_adv = NaN
        def _transcriber(self, x):
            return torch.sigmoid(self.frame.bias).expand(x.shape[0], x.shape[2], -1) + 0.0 * x.sum(), None

    m = Dead("unet").to(dev)
    m.transcriber = m._transcriber
    x = _spec_like(1, 4).to(dev)
    vat = R.VAT.UNet_VAT(1e-6, 2.0, 1, False, strict=False)
    vat(m, x)                                  # deferred mode: the flag is raised on the device, not synchronised here
    with pytest.raises(AssertionError, match="r_adv has nan"):
        vat.check()
    default = R.VAT.UNet_VAT(1e-6, 2.0, 1, False)          # as the reference: raised before the loss is handed back
    assert default.strict
    with pytest.raises(AssertionError, match="please debug tune down the XI"):
        default(m, x)
    # flavours whose reference has no assert (model/VAT.py, self_attention_VAT.stepwise_VAT) carry on with NaN
    m.forward = lambda z: m._transcriber(z)
    loss, r_adv, _ = R.VAT.stepwise_VAT(1e-6, 2.0, 1, False)(m, x)
    assert torch.isnan(r_adv).any()


def test_nan_posterior_gives_nan_loss_and_gradient(R, dev):
    """ATen's max / clamp propagate NaN: a diverged network must show up in the logged VAT loss."""
    p = torch.rand(4, 640, 88, device=dev) * 0.98 + 0.01
    y = torch.rand(4, 640, 88, device=dev) * 0.98 + 0.01
    for fn in (R.VAT.bce_mean, R.VAT.binary_kl_div, R.VAT.mse_mean):
        assert torch.isfinite(fn(p, y))
        q = p.clone()
        q[1, 7, 3] = float("nan")
        q.requires_grad_(True)
        loss = fn(q, y)
        assert torch.isnan(loss), fn.__name__
        loss.backward()
        assert torch.isnan(q.grad[1, 7, 3]) and torch.isfinite(q.grad[0]).all()


def test_vat_rejects_what_it_does_not_implement(R, dev):
    with pytest.raises(NotImplementedError):
        R.VAT.UNet_VAT(1e-6, 2.0, 2, False)
    with pytest.raises(NotImplementedError):
        R.VAT.UNet_VAT_onset(1e-6, 2.0, 1, True)      # the reference raises NameError there (UNet_onset.py:133)
    with pytest.raises(NotImplementedError):
        R.VAT.Seg_VAT(1e-6, 2.0, 1, True)             # ... and there (Segmentation.py:55)
    with pytest.raises(ValueError):
        R.VAT.stepwise_VAT_frame_stack(1e-6, 2.0, 1, "onset")
    vat = R.VAT.UNet_VAT(1e-6, 2.0, 1, False)
    with pytest.raises(R._lib.RvbError):
        vat(None, torch.zeros(1, 1, 4, 229))   # CPU tensor: no fallback


def test_l2_normalize_and_n_power_zero(R, dev):
    from oracle import vat as OV
    torch.manual_seed(0)
    d = torch.randn(3, 1, 11, 229)
    out = R.VAT.l2_normalize(d.to(dev)).cpu()
    assert float((out - OV.l2_normalize(d)).abs().max()) < 1e-6
    from reconvat_b200.standin import StandInTranscriber
    m = StandInTranscriber("unet").to(dev)
    x = _spec_like(2, 6).to(dev)
    vat = R.VAT.UNet_VAT(1e-6, 2.0, 0, False)
    loss, r_adv, d_hat = vat(m, x)
    assert torch.allclose(r_adv.norm(dim=-1), torch.full_like(r_adv[..., 0], 2.0), rtol=1e-5)
    assert torch.allclose(r_adv, 2.0 * d_hat)


@pytest.mark.parametrize("shape,warm", [((32, 1, 640, 229), 0), ((8, 640, 229), 3), ((2, 1, 33, 229), 1), ((3, 1, 7, 100), 0),
                                        ((1, 1, 1324, 229), 2)])
def test_in_kernel_direction_draw_is_torch_randn_like(R, dev, shape, warm):
    """a11 (model/self_attention_VAT.py:172): rvb_vat_perturb_draw draws d inside the kernel -- Philox4x32-10 +
    Box-Muller with ATen's element <-> (thread, call, component) mapping.  Same seed and generator position in, the
    SAME BITS as torch.randn_like out, the generator advanced by the same amount, x_adv as rvb_vat_perturb on that d.
    Shapes: the benchmark tensor (16 Philox bands), the O&F 3-D input, a tensor smaller than ATen's grid, a row length
    that is not 229, and one whose rows straddle a band boundary."""
    from reconvat_b200 import VAT
    gen = torch.cuda.default_generators[dev.index or 0]
    torch.manual_seed(1234 + warm)
    for _ in range(warm):
        torch.randn(1000 + 77 * warm, device=dev)            # the draw does not start at offset 0
    x = torch.rand(shape, device=dev)
    state = gen.get_state()
    d_ref = torch.randn_like(x)
    after = gen.get_offset()
    gen.set_state(state)
    n_rows, row_len = x.numel() // x.shape[-1], x.shape[-1]
    d, x_adv = torch.full_like(x, float("nan")), torch.empty_like(x)
    VAT._perturb_draw(x, x_adv, d, n_rows, row_len, 0.1, True)
    assert gen.get_offset() == after
    assert torch.equal(d, d_ref)
    gen.set_state(state)
    assert torch.equal(VAT.randn_like(x), d_ref) and gen.get_offset() == after      # the stand-alone kernel
    want = torch.empty_like(x)
    R._lib.call("rvb_vat_perturb", x.data_ptr(), d_ref.data_ptr(), want.data_ptr(), n_rows, row_len, 0.1, 1)
    assert torch.equal(x_adv, want)
    nxt = torch.randn(5, device=dev)                         # ... and the stream continues where ATen's would
    gen.set_state(state)
    torch.randn_like(x)
    assert torch.equal(nxt, torch.randn(5, device=dev))


def test_module_with_fused_draw_equals_module_with_aten_draw(R, dev, monkeypatch):
    """The whole module, same seed: d drawn in the kernel vs by torch.randn_like -- identical r_adv, d_hat and loss
    bits; and a device-resident stream (CUDA-graph mode) that advances from call to call."""
    from reconvat_b200 import VAT
    from reconvat_b200.standin import StandInTranscriber
    m = StandInTranscriber("unet", seed=4).to(dev)
    x = _spec_like(3, 40).to(dev)
    vat = R.VAT.UNet_VAT(0.1, 2.0, 1, False)
    torch.manual_seed(8)
    a = vat(m, x)
    for mode in ("aten", "fused"):
        monkeypatch.setenv("RVB_DRAW", mode)
        torch.manual_seed(8)
        b = vat(m, x)
        assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and float(a[0]) == float(b[0]), mode
    monkeypatch.delenv("RVB_DRAW")
    sc = VAT.Scratch(dev)
    n_rows = x.numel() // 229
    outs = []
    for _ in range(3):
        d, xa = torch.empty_like(x), torch.empty_like(x)
        VAT._perturb_draw(x, xa, d, n_rows, 229, 0.1, True, rng_state=sc.rng_state)
        outs.append(d)
    st = sc.rng_state.cpu()
    _, inc = VAT.philox_geometry(x.numel(), dev)
    assert int(st[2]) == 0 and not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    # replay i of the device stream == torch.randn_like at offset0 + i * increment
    gen = torch.cuda.default_generators[dev.index or 0]
    saved = gen.get_state()
    gen.set_offset(int(st[1]) - inc)
    assert torch.equal(torch.randn_like(x), outs[2])
    gen.set_state(saved)
