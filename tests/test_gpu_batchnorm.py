"""reconvat_b200.batchnorm.BatchNorm2d (consumer side: the U-Net's train-mode BatchNorm) against
torch.nn.BatchNorm2d on the same GPU and against float64."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

# (N, C, H, W): the U-Net's own tensors at B = 2 (self_attention_VAT.py:890-893), ragged sizes (scalar path: hw % 4 != 0),
# one sample, one channel, tiny planes
SHAPES = [(2, 16, 640, 229), (2, 32, 320, 114), (2, 64, 160, 57), (2, 128, 80, 28), (3, 5, 17, 13), (1, 7, 33, 9),
          (4, 1, 50, 50), (5, 3, 1, 2)]


def _pair(c, dev, **kw):
    from reconvat_b200 import batchnorm
    torch.manual_seed(c)
    ref = nn.BatchNorm2d(c, **kw).to(dev)
    if ref.affine:
        with torch.no_grad():
            ref.weight.uniform_(0.5, 1.5)
            ref.bias.uniform_(-0.5, 0.5)
    ours = batchnorm.convert(copy.deepcopy(ref))
    assert type(ours) is batchnorm.BatchNorm2d and type(ref) is nn.BatchNorm2d
    return ref, ours


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", SHAPES)
def test_train_forward_backward_match_torch_and_float64(shape):
    dev = torch.device("cuda:0")
    n, c, h, w = shape
    ref, ours = _pair(c, dev, momentum=0.1)
    g = torch.Generator().manual_seed(sum(shape))
    # activations with a mean far from zero and a per-channel scale: E[x^2] - mean^2 would cancel
    x = (torch.randn(shape, generator=g) * torch.rand(1, c, 1, 1, generator=g).mul(3).add(0.1) + 5.0).to(dev)
    dy = torch.randn(shape, generator=g).to(dev)
    outs = []
    for m, dt in ((ref, torch.float32), (ours, torch.float32), (copy.deepcopy(ref).double(), torch.float64)):
        for step in range(2):                                   # two steps: the running statistics accumulate
            xi = x.to(dt).clone().requires_grad_(True)
            m.zero_grad()
            y = m(xi)
            y.backward(dy.to(dt))
        outs.append((y.detach(), xi.grad, m.weight.grad, m.bias.grad, m.running_mean, m.running_var, m.num_batches_tracked))
    (y_r, dx_r, dg_r, db_r, rm_r, rv_r, nb_r), (y_o, dx_o, dg_o, db_o, rm_o, rv_o, nb_o), (y_t, dx_t, dg_t, db_t, rm_t, rv_t, _) = outs
    assert int(nb_o) == int(nb_r) == 2
    # against torch (cuDNN / native) on the same GPU
    assert _rel(y_o, y_r) < 2e-6 and _rel(dx_o, dx_r) < 2e-5
    assert _rel(dg_o, dg_r) < 2e-5 and _rel(db_o, db_r) < 2e-5
    assert _rel(rm_o, rm_r) < 1e-6 and _rel(rv_o, rv_r) < 1e-5
    # against float64: at least as close as torch is (float64 partial sums, shifted variance)
    for ours_v, torch_v, truth in ((y_o, y_r, y_t), (dx_o, dx_r, dx_t), (dg_o, dg_r, dg_t), (rv_o, rv_r, rv_t)):
        assert _rel(ours_v, truth) <= max(2.0 * _rel(torch_v, truth), 1e-6)


@pytest.mark.parametrize("shape", [(2, 16, 640, 229), (2, 32, 320, 114), (2, 48, 160, 57), (2, 128, 80, 28), (3, 8, 17, 13),
                                   (1, 4, 3, 5), (2, 1024, 6, 5), (2, 12, 30, 7)])
def test_channels_last_kernels_match_torch_and_float64(shape, monkeypatch):
    """torch.channels_last tensors through the rvb NHWC kernels (RVB_BN_NHWC=1; no layout conversion: the output and
    the input gradient come back channels_last) -- training and eval, against torch on the same layout and float64."""
    from reconvat_b200 import _lib, batchnorm
    monkeypatch.setenv("RVB_BN_NHWC", "1")
    launches0 = _lib.launch_count()
    dev = torch.device("cuda:0")
    n, c, h, w = shape
    ref, ours = _pair(c, dev, momentum=0.1)
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * torch.rand(1, c, 1, 1, generator=g).mul(3).add(0.1) - 4.0).to(dev)
    x = x.contiguous(memory_format=torch.channels_last)
    dy = torch.randn(shape, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    assert batchnorm._is_nhwc(x) == (c % 4 == 0 and h * w > 1)
    res = []
    for m, dt in ((ref, torch.float32), (ours, torch.float32), (copy.deepcopy(ref).double(), torch.float64)):
        for step in range(2):
            xi = x.to(dt).clone(memory_format=torch.preserve_format).requires_grad_(True)
            m.zero_grad()
            y = m(xi)
            y.backward(dy.to(dt))
        m.eval()
        ye = m(x.to(dt)).detach()
        m.train()
        res.append((y.detach(), xi.grad, m.weight.grad, m.bias.grad, m.running_mean, m.running_var, ye))
    r, o, t = res
    assert _lib.launch_count() > launches0                      # the rvb kernels ran (not cuDNN)
    if batchnorm._is_nhwc(x):
        assert o[0].is_contiguous(memory_format=torch.channels_last) and o[1].is_contiguous(memory_format=torch.channels_last)
        assert o[6].is_contiguous(memory_format=torch.channels_last)
    for i, tol in ((0, 2e-6), (1, 2e-5), (2, 2e-5), (3, 2e-5), (4, 1e-6), (5, 1e-5), (6, 2e-6)):
        assert _rel(o[i], r[i]) < tol, (i, _rel(o[i], r[i]))
    for i in (0, 1, 2, 5):
        assert _rel(o[i], t[i]) <= max(2.0 * _rel(r[i], t[i]), 1e-6), i


def test_eval_mode_affine_off_untracked_and_errors():
    from reconvat_b200 import _lib, batchnorm
    dev = torch.device("cuda:0")
    x = torch.randn(3, 6, 20, 12, device=dev) * 2 + 1
    dy = torch.randn_like(x)
    # eval mode: running statistics, gradients flow through the affine map only
    ref, ours = _pair(6, dev)
    for m in (ref, ours):
        m(x)                                                    # one training step so that the statistics are not 0 / 1
        m.eval()
    res = []
    for m in (ref, ours):
        xi = x.clone().requires_grad_(True)
        m.zero_grad()
        y = m(xi)
        y.backward(dy)
        res.append((y.detach(), xi.grad, m.weight.grad.clone(), m.bias.grad.clone()))
    for a, b in zip(res[1], res[0]):
        assert _rel(a, b) < 2e-6
    assert torch.equal(ours.running_mean, ref.running_mean) or _rel(ours.running_mean, ref.running_mean) < 1e-6
    # affine=False, track_running_stats=False (batch statistics in eval mode too), momentum=None (cumulative average)
    for kw in (dict(affine=False), dict(track_running_stats=False), dict(momentum=None)):
        ref, ours = _pair(6, dev, **kw)
        for mode in ("train", "eval"):
            getattr(ref, mode)(); getattr(ours, mode)()
            xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
            ya, yb = ref(xa), ours(xb)
            ya.backward(dy); yb.backward(dy)
            assert _rel(yb, ya) < 2e-6 and _rel(xb.grad, xa.grad) < 2e-5, (kw, mode)
        if ref.running_mean is not None:
            assert _rel(ours.running_mean, ref.running_mean) < 1e-6 and _rel(ours.running_var, ref.running_var) < 1e-5
    # non-contiguous input (channels_last memory format) gives the same values
    ref, ours = _pair(6, dev)
    xc = x.contiguous(memory_format=torch.channels_last)
    assert _rel(ours(xc), ref(x)) < 2e-6
    # no CPU path; one value per channel in training mode is an error, as in torch
    with pytest.raises(_lib.RvbError):
        batchnorm.BatchNorm2d(6)(torch.zeros(2, 6, 4, 4))
    with pytest.raises(ValueError):
        ours.train()(torch.zeros(1, 6, 1, 1, device=dev))
    with pytest.raises(ValueError):
        ours(torch.zeros(2, 6, 4, device=dev))                   # 3-D input: _check_input_dim, as nn.BatchNorm2d


def test_state_dict_interchange_and_graph_capture():
    from reconvat_b200 import batchnorm
    dev = torch.device("cuda:0")
    ref, ours = _pair(16, dev)
    assert sorted(ours.state_dict()) == sorted(ref.state_dict())
    ours.load_state_dict(ref.state_dict(), strict=True)
    net = nn.Sequential(nn.Conv2d(1, 16, 3, padding=1), nn.BatchNorm2d(16), nn.LeakyReLU()).to(dev)
    conv = batchnorm.convert(copy.deepcopy(net))
    assert type(conv[1]) is batchnorm.BatchNorm2d and type(net[1]) is nn.BatchNorm2d
    x = torch.randn(2, 1, 64, 48, device=dev)
    assert _rel(conv(x), net(x)) < 2e-6
    # CUDA-graph capture: no host synchronisation, no allocation outside the capture pool
    static_x = x.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        conv(static_x)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_y = conv(static_x)
    static_x.copy_(torch.randn_like(x))
    graph.replay()
    torch.cuda.synchronize()
    net[1].load_state_dict(conv[1].state_dict())
    net[1].momentum = 0.0                                        # compare the outputs, not another statistics update
    assert _rel(static_y, net(static_x)) < 2e-6


def test_reference_unet_with_the_batchnorm_seam():
    """The reference's own UNet built behind install(attention=True) and behind install(attention=True,
    batchnorm=True): same parameters (same seed), same batch -> same losses and parameter gradients."""
    import refmodels as RM
    from oracle import reference_loader as RL
    from reconvat_b200 import batchnorm
    if not RL.available():
        pytest.skip("no reference snapshot")
    dev = torch.device("cuda:0")
    a = RL.load_patched(attention=True)
    b = RL.load_patched(attention=True, batchnorm=True)
    ma, mb = RM.build(a, "unet", dev, 1e-6, 2.0), RM.build(b, "unet", dev, 1e-6, 2.0)
    n_ours = sum(isinstance(m, batchnorm.BatchNorm2d) for m in mb.modules())
    assert n_ours == 30 and not any(isinstance(m, batchnorm.BatchNorm2d) for m in ma.modules())
    assert all(torch.equal(p, q) for p, q in zip(ma.state_dict().values(), mb.state_dict().values()))
    batch = RM.batch(2, 3, dev)

    def run(m, cudnn=True):
        m.train()
        m.zero_grad()
        with torch.backends.cudnn.flags(enabled=cudnn):
            _, losses, _ = m.run_on_batch(batch, None, False)
            sum(losses.values()).backward()
        gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None))
        return {k: float(v.detach()) for k, v in losses.items()}, float(gn)
    sd0 = copy.deepcopy(ma.state_dict())
    l_cudnn, g_cudnn = run(ma)
    sd_cudnn = copy.deepcopy(ma.state_dict())
    ma.load_state_dict(sd0)
    l_native, g_native = run(ma, cudnn=False)                    # the yardstick: torch's OTHER BatchNorm (and conv) kernels
    l_ours, g_ours = run(mb)
    for k in l_cudnn:
        assert abs(l_ours[k] - l_cudnn[k]) <= 5e-4 * abs(l_cudnn[k]) + 1e-7, (k, l_ours[k], l_cudnn[k])   # measured 5e-5
    # the gradient of this randomly initialised network is dominated by rounding noise (every convolution bias in front
    # of a BatchNorm has a zero true gradient): the reference differs from itself by 2 % in |g| between its cuDNN and its
    # native kernels; ours must lie as close
    yard = abs(g_native - g_cudnn) / g_cudnn
    assert abs(g_ours - g_cudnn) / g_cudnn <= max(5e-3, 2.0 * yard), (g_ours, g_cudnn, g_native)
    # running statistics after the step (deterministic functions of the forward activations)
    for (k, p), q in zip(sd_cudnn.items(), mb.state_dict().values()):
        if "running_" in k:
            assert _rel(q, p) < 1e-3, k
