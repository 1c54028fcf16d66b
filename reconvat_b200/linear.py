"""Bias-free linear projections on tcgen05 tensor cores in 3xTF32.

The sequence model of the ReconVAT U-Net projects every frame three times (``W_q``, ``W_k``, ``W_v`` of
``MutliHeadAttention1D``, model/self_attention_VAT.py:54-56, 70-71): 229 -> 916 on B * 640 rows, forward and backward
77 GFLOP per call at B = 32.  PyTorch runs ``nn.Linear`` in true fp32 (``torch.backends.cuda.matmul.allow_tf32`` is off
by default), i.e. on the SIMT pipes.  ``projections(x, weights)`` computes the same products as three-pass split-TF32
contractions on the tensor cores (``rvb_gemm_nt_tf32x3``: hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM, relative
error ~2^-21 -- the size of fp32 SGEMM's own accumulation rounding), with an autograd backward built from the same
kernel:

    y_i  = x   . w_i^T                    A = x planes,                B = w_i planes            (contraction over K)
    dx   = sum_i dy_i . w_i               A = [dy_1 | dy_2 | ..] planes, B = [w_1; w_2; ..]^T    (over sum N_i)
    dw_i = dy_i^T . x                     A = dy_i^T planes,           B = x^T planes            (over the rows)

Operand planes (tf32 hi / lo, contraction length padded to a multiple of 32) come from ``rvb_split_tf32``, which also
transposes where the contraction runs over rows.
"""
import torch

from . import _lib


def _pad32(n):
    return (n + 31) // 32 * 32


def _planes(rows, width, device):
    return (torch.empty((rows, width), dtype=torch.float32, device=device),
            torch.empty((rows, width), dtype=torch.float32, device=device))


def _split(x, planes=None, transpose=False, offset=0, zero_to=None):
    """x: 2-D float32 CUDA tensor with unit column stride.  Returns (hi, lo)."""
    rows, cols = x.shape
    extent = rows if transpose else cols
    if planes is None:
        width = _pad32(extent)
        planes = _planes(cols if transpose else rows, width, x.device)
        zero_to = width
    hi, lo = planes
    zero_to = offset + extent if zero_to is None else zero_to
    _lib.call("rvb_split_tf32", _lib.ptr(x), rows, cols, x.stride(0), int(transpose), hi.data_ptr(), lo.data_ptr(),
              hi.shape[1], offset, zero_to)
    return planes


_N_SM = {}


def _sm_count(device):
    index = device.index if device.index is not None else torch.cuda.current_device()
    if index not in _N_SM:
        _N_SM[index] = torch.cuda.get_device_properties(index).multi_processor_count
    return _N_SM[index]


def _gemm_nt(a, b, m, n, out):
    """out[m][n] = A . B^T from (hi, lo) planes with the same padded contraction length.  Products with fewer output
    tiles (128 x 256) than SMs and a long contraction are cut along it: every slice writes its own partial product and
    the partials are added in a fixed order."""
    k_pad = a[0].shape[1]
    assert b[0].shape[1] == k_pad and out.stride(1) == 1
    tiles = -(-m // 128) * -(-n // 256)
    blocks = k_pad // 32
    sms = _sm_count(out.device)
    k_split = 1
    if tiles < 4 * sms and blocks >= 16:
        # fewer than four rounds of tiles: cut the contraction until there are ~six (no tail of half-empty rounds),
        # slices of at least eight 32-element blocks
        k_split = max(1, min(blocks // 8, -(-6 * sms // tiles)))
        per = -(-blocks // k_split)
        k_split = -(-blocks // per)                           # no empty slice
    if k_split == 1:
        _lib.call("rvb_gemm_nt_tf32x3", a[0].data_ptr(), a[1].data_ptr(), m, b[0].data_ptr(), b[1].data_ptr(), n, k_pad,
                  out.data_ptr(), out.stride(0), 1, 0)
        return out
    part = torch.empty((k_split, m, n), dtype=torch.float32, device=out.device)
    _lib.call("rvb_gemm_nt_tf32x3", a[0].data_ptr(), a[1].data_ptr(), m, b[0].data_ptr(), b[1].data_ptr(), n, k_pad,
              part.data_ptr(), n, k_split, m * n)
    torch.sum(part, dim=0, out=out)
    return out


class _Projections(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, *weights):
        for t in (x,) + weights:
            if not t.is_cuda or t.dtype != torch.float32:
                raise _lib.RvbError("reconvat_b200 projections need CUDA float32 tensors (got %s, %s); there is no CPU "
                                    "path" % (t.device, t.dtype))
        x = x.contiguous()
        m, k = x.shape
        xa = _split(x)
        outs = []
        for w in weights:
            w = w.contiguous()
            assert w.shape[1] == k
            outs.append(_gemm_nt(xa, _split(w), m, w.shape[0], torch.empty((m, w.shape[0]), dtype=torch.float32, device=x.device)))
        ctx.save_for_backward(x, *weights)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dys):
        x, *weights = ctx.saved_tensors
        m, k = x.shape
        dev = x.device
        dys = [None if d is None else d.contiguous() for d in dys]
        live = [(d, w) for d, w in zip(dys, weights) if d is not None]
        dx = None
        if ctx.needs_input_grad[0] and live:
            n_tot = sum(w.shape[0] for _, w in live)
            width = _pad32(n_tot)
            da, wt = _planes(m, width, dev), _planes(k, width, dev)
            off = 0
            for i, (d, w) in enumerate(live):
                last = i == len(live) - 1
                _split(d, da, False, off, width if last else None)             # [dy_1 | dy_2 | ...]
                _split(w.contiguous(), wt, True, off, width if last else None)  # [w_1; w_2; ...]^T
                off += w.shape[0]
            dx = _gemm_nt(da, wt, m, k, torch.empty((m, k), dtype=torch.float32, device=dev))
            del da, wt
        # dW_i = dy_i^T . x for all i in ONE contraction: A = [dy_1 | dy_2 | ...]^T stacked along the rows (each dy_i^T
        # written by a transposing split into its row block), B = x^T; the contraction runs over the m rows
        want = [i for i, d in enumerate(dys) if d is not None and ctx.needs_input_grad[1 + i]]
        dws = [None] * len(weights)
        if want:
            n_rows = sum(weights[i].shape[0] for i in want)
            m_pad = _pad32(m)
            dt = _planes(n_rows, m_pad, dev)
            xt = _split(x, transpose=True)                                       # (k, m_pad)
            off = 0
            for i in want:
                n_i = weights[i].shape[0]
                _split(dys[i], (dt[0][off:off + n_i], dt[1][off:off + n_i]), True, 0, m_pad)
                off += n_i
            dw_cat = _gemm_nt(dt, xt, n_rows, k, torch.empty((n_rows, k), dtype=torch.float32, device=dev))
            off = 0
            for i in want:
                n_i = weights[i].shape[0]
                dws[i] = dw_cat[off:off + n_i]
                off += n_i
            del dt, xt
        return (dx,) + tuple(dws)


def projections(x, weights):
    """``[F.linear(x, w) for w in weights]`` for a (..., K) input and (N_i, K) weights, on the tensor cores (see the
    module docstring); differentiable w.r.t. x and every weight."""
    lead = x.shape[:-1]
    outs = _Projections.apply(x.reshape(-1, x.shape[-1]), *weights)
    return [o.view(*lead, o.shape[-1]) for o in outs]
