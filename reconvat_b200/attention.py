"""Drop-in for the reference's ``MutliHeadAttention1D`` (model/self_attention_VAT.py:22-91; the same class is repeated in
model/UNet_onset.py:22, model/onset_frame_VAT.py:16 and model/self_attention.py:6): 1-D local-window multi-head
attention with a relative position term, the sequence model of the ReconVAT U-Net (Spec2Roll.lstm1 / Roll2Spec.lstm2,
:934,:953).  SURVEY.md 8f row f2: it is the CALLER of the hot path, and the worst memory amplifier in it -- the
reference unfolds k and v to (B, L, C, W), 73 MB per 20 s segment each at C = 916, W = 31, and autograd keeps both.

Same constructor, parameters (``W_q``, ``W_k``, ``W_v``, ``rel`` -> state_dict compatible) and return values
``(out (B, L, C), attention (B, L, groups, W))``.  The three projections keep their ``nn.Linear`` parameters but run as
3xTF32 tcgen05 contractions (``reconvat_b200.linear.projections``: PyTorch's fp32 SGEMMs were 61 % of the layer;
``RVB_ATTN_PROJ=torch`` restores ``nn.Linear``); everything after them is one kernel forward and two backward
(librvb.so, rvb_attention.cu); the relative-position terms (q . rel, dE . rel^T, d rel) do not involve k and are small
batched GEMMs (torch / cuBLAS).
Scope cuts, raising: ``stride != 1`` and ``bias=True`` (no reference model uses them; with a bias the zero padding
rows would become the bias vector).
"""
import os

import torch
import torch.nn as nn
import torch.nn.init as init

from . import _lib, linear


def _projection_path(x):
    """'tc' (3xTF32 tcgen05 contractions, reconvat_b200.linear) or 'torch' (nn.Linear / einsum on cuBLAS SGEMM).
    RVB_ATTN_PROJ forces one; by default the tensor-core path takes over from 16 384 rows (B >= 26 segments of 640
    frames): below that its operand-split passes and extra launches cost more than the SIMT GEMMs they replace
    (B = 8: 1.8 vs 1.3 ms per layer call; B = 32: 3.3 vs 4.4 ms, profiles/r02_attention.txt)."""
    forced = os.environ.get("RVB_ATTN_PROJ")
    if forced in ("tc", "torch"):
        return forced
    return "tc" if x.numel() // x.shape[-1] >= 16384 else "torch"


class _LocalAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, rel, groups, window):
        for t in (q, k, v):
            if not t.is_cuda or t.dtype != torch.float32:
                raise _lib.RvbError("reconvat_b200 attention needs CUDA float32 tensors (got %s, %s); there is no CPU "
                                    "path" % (t.device, t.dtype))
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        B, L, C = q.shape
        D = C // groups
        bias = None
        if rel is not None:
            # q . rel does not involve k: one batched GEMM over the heads (cuBLAS), added to the energies in the kernel
            bias = torch.einsum("nhc,hcw->nhw", q.view(B * L, groups, D), rel.view(groups, D, window)).contiguous()
        out = torch.empty_like(q)
        att = torch.empty((B, L, groups, window), dtype=torch.float32, device=q.device)
        _lib.call("rvb_local_attn_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), None if bias is None else bias.data_ptr(),
                  B, L, groups, D, window, out.data_ptr(), att.data_ptr())
        ctx.save_for_backward(q, k, v, rel, att)
        ctx.dims = (B, L, groups, D, window)
        ctx.mark_non_differentiable(att)          # the reference only plots it (model/self_attention_VAT.py:938)
        return out, att

    @staticmethod
    def backward(ctx, dout, _datt):
        q, k, v, rel, att = ctx.saved_tensors
        B, L, G, D, W = ctx.dims
        dout = dout.contiguous()
        dE = torch.empty_like(att)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        _lib.call("rvb_local_attn_bwd_q", dout.data_ptr(), att.data_ptr(), k.data_ptr(), v.data_ptr(), B, L, G, D, W,
                  dE.data_ptr(), dq.data_ptr())
        _lib.call("rvb_local_attn_bwd_kv", q.data_ptr(), dout.data_ptr(), att.data_ptr(), dE.data_ptr(), B, L, G, D, W,
                  dk.data_ptr(), dv.data_ptr())
        drel = None
        if rel is not None:
            dE2, rel3 = dE.view(B * L, G, W), rel.view(G, D, W)
            dq = dq + torch.einsum("nhw,hcw->nhc", dE2, rel3).reshape(B, L, G * D)      # dE . rel^T
            if _projection_path(q) == "torch":
                drel = torch.einsum("nhc,nhw->hcw", q.view(B * L, G, D), dE2).reshape(G * D, W)
            else:
                # d rel[h] = q_h^T . dE_h: 229 x 31 outputs, 20 480 terms -- cuBLAS picks a 464 us kernel for this shape;
                # here it is the split-K tensor-core contraction (operands transposed while they are split)
                q2, drel = q.view(B * L, G * D), torch.empty((G * D, W), dtype=torch.float32, device=q.device)
                for h in range(G):
                    qt = linear._split(q2[:, h * D:(h + 1) * D], transpose=True)            # (D, n_pad)
                    et = linear._split(dE2[:, h, :], transpose=True)                         # (W, n_pad)
                    linear._gemm_nt(qt, et, D, W, drel[h * D:(h + 1) * D])
        return dq, dk, dv, drel, None, None


class MutliHeadAttention1D(nn.Module):
    def __init__(self, in_features, out_features, kernel_size, stride=1, groups=1, position=True, bias=False):
        """kernel_size is the 1D local attention window size"""
        super().__init__()
        if stride != 1:
            raise NotImplementedError("reconvat_b200 MutliHeadAttention1D: stride=%r (every reference model uses 1)" % stride)
        if bias:
            raise NotImplementedError("reconvat_b200 MutliHeadAttention1D: bias=True (never used by the reference)")
        self.out_features = out_features
        self.kernel_size = kernel_size
        self.stride = stride
        self.position = position
        self.padding = (kernel_size - 1) // 2
        self.groups = groups
        assert self.out_features % self.groups == 0, (
            f"out_channels should be divided by groups. (example: out_channels: 40, groups: 4). "
            f"Now out_channels={self.out_features}, groups={self.groups}")
        assert (kernel_size - 1) % 2 == 0, "kernal size must be odd number"
        if self.position:
            self.rel = nn.Parameter(torch.randn(1, out_features, kernel_size), requires_grad=True)
        self.W_k = nn.Linear(in_features, out_features, bias=bias)
        self.W_q = nn.Linear(in_features, out_features, bias=bias)
        self.W_v = nn.Linear(in_features, out_features, bias=bias)
        self.reset_parameters()

    def forward(self, x):
        if _projection_path(x) == "torch":
            q, k, v = self.W_q(x), self.W_k(x), self.W_v(x)              # zero padding rows project to zero (no bias)
        else:
            q, k, v = linear.projections(x, [self.W_q.weight, self.W_k.weight, self.W_v.weight])
        return _LocalAttention.apply(q, k, v, self.rel[0] if self.position else None, self.groups, self.kernel_size)

    def reset_parameters(self):
        init.kaiming_normal_(self.W_k.weight, mode='fan_out', nonlinearity='relu')
        init.kaiming_normal_(self.W_v.weight, mode='fan_out', nonlinearity='relu')
        init.kaiming_normal_(self.W_q.weight, mode='fan_out', nonlinearity='relu')
        if self.position:
            init.normal_(self.rel, 0, 1)
