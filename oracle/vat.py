"""Oracle restatement of the VAT perturbation loop (CPU torch, fp32 / fp64).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Paths relative to /root/reference.

Two forms are given and tested against each other and against the reference:

* ``*_autograd``  -- the reference's op sequence verbatim through torch autograd
  (self_attention_VAT.py:162-202, UNet_onset.py:116-162,
  onset_frame_VAT.py:175-207, VAT.py:20-40);
* closed form     -- the same arithmetic with the autograd chain
  clamp <- add <- mul <- div <- norm written out, which is what the CUDA
  kernels implement:
      n      = ||d||_row ;  dhat = d / n ;  s = x + XI*dhat
      m      = [0 <= s <= 1]                      (clamp passes grad on the closed interval)
      gm     = g * m                              (g = dL/dx_adv from the model)
      d.grad = XI*gm/n - d * (sum_row(XI*gm*d) / n^3)
      d'     = d.grad * scale                     (scale = 1e10 at :184)
      r_adv  = eps * d'/||d'||_row                (:188)
"""
import torch
import torch.nn.functional as F


def l2_normalize(d):
    """self_attention_VAT.py:240-246 (binwise=False): per last-axis row."""
    return d / torch.norm(d, dim=-1, keepdim=True)


def perturb(x, d, xi, clamp=True):
    """self_attention_VAT.py:176-177  (VAT.py:27-28 has no clamp)."""
    s = x + xi * l2_normalize(d)
    return s.clamp(0, 1) if clamp else s


def power_grad_closed_form(x, d, g, xi, scale=1e10, clamp=True):
    """d.grad * scale, given g = dL/dx_adv (self_attention_VAT.py:183-184)."""
    n = torch.norm(d, dim=-1, keepdim=True)
    dhat = d / n
    if clamp:
        s = x + xi * dhat
        m = ((s >= 0) & (s <= 1)).to(g.dtype)
        gm = g * m
    else:
        gm = g
    gd = xi * gm
    grad = gd / n - d * ((gd * d).sum(-1, keepdim=True) / (n * n * n))
    return grad * scale


def power_grad_autograd(x, d, g, xi, scale=1e10, clamp=True):
    d = d.clone().requires_grad_(True)
    x_adv = perturb(x, d, xi, clamp)
    x_adv.backward(g)
    return d.grad.detach() * scale


def finalize(x, dprime, eps, clamp=True):
    """self_attention_VAT.py:188,194,202 -> r_adv, x_adv, dhat'."""
    dhat = l2_normalize(dprime)
    r_adv = eps * dhat
    s = x + r_adv
    return r_adv, (s.clamp(0, 1) if clamp else s), dhat


def bce_mean(p, y):
    """F.binary_cross_entropy(p, y) (mean; log clamped at -100) :182,200."""
    return F.binary_cross_entropy(p, y)


def bce_mean_grad(p, y):
    """d mean-BCE / d p, ATen's formula: (p - y) / max((1-p)*p, 1e-12) / numel."""
    return (p - y) / torch.clamp((1 - p) * p, min=1e-12) / p.numel()


def binary_kl_div(y_pred, y_ref, lo=1e-4):
    """self_attention_VAT.py:248-255 (lo=1e-4) / onset_frame_VAT.py:151-156 (lo=0)."""
    y_pred = torch.clamp(y_pred, lo, 0.9999)
    y_ref = torch.clamp(y_ref, lo, 0.9999)
    q = torch.stack((y_pred, 1 - y_pred), -1)
    p = torch.stack((y_ref, 1 - y_ref), -1)
    return F.kl_div(p.log(), q, reduction="batchmean")


def vat_unet(transcribe, x, d, xi, eps, scale=1e10, clamp=True):
    """Whole UNet_VAT.forward (self_attention_VAT.py:162-202) with ``d``
    injected instead of drawn; ``transcribe(x) -> y`` is the model's frame
    posterior.  Returns (vat_loss, r_adv, dhat', g)."""
    with torch.no_grad():
        y_ref = transcribe(x)
    x_adv = perturb(x, d, xi, clamp).detach().requires_grad_(True)
    loss = bce_mean(transcribe(x_adv), y_ref)
    (g,) = torch.autograd.grad(loss, x_adv)
    dprime = power_grad_closed_form(x, d, g, xi, scale, clamp)
    r_adv, x_adv2, dhat = finalize(x, dprime, eps, clamp)
    vat_loss = bce_mean(transcribe(x_adv2), y_ref)
    return vat_loss, r_adv, dhat, g
