"""Oracle restatement of the reference's local-window attention (model/self_attention_VAT.py:61-88).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain torch on CPU, the reference's op sequence: F.pad, the three
projections, unfold, + rel, (q*k).sum, softmax, (att*v).sum -- including the (B, L, C, W) unfolded tensors the CUDA
kernels avoid."""
import torch
import torch.nn.functional as F


def local_attention(x, w_q, w_k, w_v, rel, groups, kernel_size):
    """x (B, L, F); w_* (C, F) nn.Linear weights (no bias); rel (1, C, W) or None.
    Returns (out (B, L, C), attention (B, L, groups, W))."""
    batch, seq_len, _ = x.shape
    pad = (kernel_size - 1) // 2
    C = w_q.shape[0]
    padded_x = F.pad(x, [0, 0, pad, pad])                                  # :64
    q_out = x @ w_q.t()                                                    # :65-67
    k_out = (padded_x @ w_k.t()).unfold(1, kernel_size, 1)                 # :69  (B, L, C, W)
    v_out = (padded_x @ w_v.t()).unfold(1, kernel_size, 1)                 # :72
    if rel is not None:
        k_out = k_out + rel                                                # :76
    k_out = k_out.contiguous().view(batch, seq_len, groups, C // groups, -1)
    v_out = v_out.contiguous().view(batch, seq_len, groups, C // groups, -1)
    q_out = q_out.view(batch, seq_len, groups, C // groups, 1)
    energy = (q_out * k_out).sum(-2, keepdim=True)                         # :86
    attention = F.softmax(energy, dim=-1)                                  # :88
    out = attention * v_out
    return out.sum(-1).flatten(2), attention.squeeze(3)
