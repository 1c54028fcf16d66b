"""Times the fused local-window attention (forward + backward) against the reference's op sequence run by PyTorch on
the same GPU (F.pad / unfold / softmax, model/self_attention_VAT.py:61-88, restated inline), at the U-Net's shape.
Usage: python tools/attn_bench.py [B]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reconvat_b200.attention import MutliHeadAttention1D      # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L, FIN, C, W, G = 640, 229, 916, 31, 4
dev = torch.device("cuda:0")


def unfold_attention(m, x):
    """The reference forward with m's parameters (eager PyTorch: materialises (B, L, C, W) k and v)."""
    pad = (W - 1) // 2
    px = F.pad(x, [0, 0, pad, pad])
    q = m.W_q(x).view(B, L, G, C // G, 1)
    k = (m.W_k(px).unfold(1, W, 1) + m.rel).contiguous().view(B, L, G, C // G, -1)
    v = m.W_v(px).unfold(1, W, 1).contiguous().view(B, L, G, C // G, -1)
    att = F.softmax((q * k).sum(-2, keepdim=True), dim=-1)
    return (att * v).sum(-1).flatten(2), att.squeeze(3)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (torch.cuda.max_memory_allocated() - base) / 2 ** 20


torch.manual_seed(0)
m = MutliHeadAttention1D(FIN, C, W, position=True, groups=G).to(dev)
x = torch.rand(B, L, FIN, device=dev, requires_grad=True)
go = torch.randn(B, L, C, device=dev)


def ours():
    out, _ = m(x)
    out.backward(go)


def ref():
    out, _ = unfold_attention(m, x)
    out.backward(go)


t_o, mem_o = timed(ours)
t_r, mem_r = timed(ref)
with torch.no_grad():
    o1, a1 = m(x)
    o2, a2 = unfold_attention(m, x)
print("B=%d L=%d %d->%d W=%d heads=%d" % (B, L, FIN, C, W, G))
print("fused kernels : %.3f ms fwd+bwd, peak extra memory %.0f MiB" % (t_o, mem_o))
print("eager unfold  : %.3f ms fwd+bwd, peak extra memory %.0f MiB" % (t_r, mem_r))
print("speed-up %.1fx, memory %.1fx smaller; max |out diff| / max |out| = %.2e, max |att diff| = %.2e"
      % (t_r / t_o, mem_r / max(mem_o, 1e-9), float((o1 - o2).abs().max() / o2.abs().max()), float((a1 - a2).abs().max())))
