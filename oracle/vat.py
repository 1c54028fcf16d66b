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


def mse_mean(p, y):
    """F.mse_loss (onset_frame_VAT.py:232)."""
    return F.mse_loss(p, y)


def l2_normalize_binwise(d):
    """self_attention_VAT.py:242-243 (binwise=True): d / (|d| + 1e-8), no row coupling."""
    return d / (torch.abs(d) + 1e-8)


def power_grad_binwise_autograd(x, d, g, xi, scale=1.0, clamp=True):
    """d.grad * scale for binwise=True through torch autograd, as the reference obtains it."""
    d = d.clone().requires_grad_(True)
    s = x + xi * l2_normalize_binwise(d)
    (s.clamp(0, 1) if clamp else s).backward(g)
    return d.grad.detach() * scale


def power_grad_binwise_sequence(x, d, g, xi, scale=1.0, clamp=True):
    """The same gradient written out as the fp32 op sequence autograd executes -- what the CUDA kernel follows.
    Mathematically d/dd [d / (|d| + e)] = e / (|d| + e)^2 ~ 1e-8, but autograd forms it as the difference of two
    O(1) terms, go/b - go*((d/b)/b)*sgn(d) with b = |d| + e, which cancels to fp32 rounding noise: the reference's
    binwise direction IS that noise, so parity means reproducing the sequence, not the closed form."""
    b = torch.abs(d) + 1e-8
    go = g
    if clamp:
        s = x + xi * (d / b)
        go = g * ((s >= 0) & (s <= 1)).to(g.dtype)          # clamp backward
    gdn = go * xi                                             # mul backward
    grad_a = gdn / b                                          # div backward w.r.t. the numerator
    grad_b = -gdn * ((d / b) / b)                             # ... and the denominator (ATen: -grad * ((self/other)/other))
    return (grad_a + grad_b * torch.sgn(d)) * scale           # abs backward, accumulation


def finalize_binwise(x, dprime, eps, clamp=True):
    dhat = l2_normalize_binwise(dprime)
    r_adv = eps * dhat
    s = x + r_adv
    return r_adv, (s.clamp(0, 1) if clamp else s), dhat


def vat_generic(outputs, divergences, x, d, xi, eps, scale=1.0, clamp=True, binwise=False):
    """The loop shared by every flavour (self_attention_VAT.py:115-145, onset_frame_VAT.py:222-263,
    Segmentation.py:39-77) with ``d`` injected: ``outputs(x)`` -> list of posteriors, ``divergences`` -> one
    callable (p, y) -> scalar per posterior; the losses are summed in list order.
    Returns (list of final losses, r_adv, dhat', g)."""
    with torch.no_grad():
        y_ref = outputs(x)
    norm = l2_normalize_binwise if binwise else l2_normalize
    x_adv = x + xi * norm(d)
    x_adv = (x_adv.clamp(0, 1) if clamp else x_adv).detach().requires_grad_(True)
    loss = sum(f(p, y) for f, p, y in zip(divergences, outputs(x_adv), y_ref))
    (g,) = torch.autograd.grad(loss, x_adv)
    if binwise:
        dprime = power_grad_binwise_sequence(x, d, g, xi, scale, clamp)
        r_adv, x_adv2, dhat = finalize_binwise(x, dprime, eps, clamp)
    else:
        dprime = power_grad_closed_form(x, d, g, xi, scale, clamp)
        r_adv, x_adv2, dhat = finalize(x, dprime, eps, clamp)
    with torch.no_grad():
        final = [f(p, y) for f, p, y in zip(divergences, outputs(x_adv2), y_ref)]
    return final, r_adv, dhat, g


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
