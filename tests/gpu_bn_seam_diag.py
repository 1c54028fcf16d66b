"""Diagnostic (GPU): parameter gradients of the reference UNet with cuDNN BN, native BN (cudnn off) and rvb BN."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refmodels as RM
from oracle import reference_loader as RL
dev = torch.device("cuda:0")
a = RL.load_patched(attention=True)
b = RL.load_patched(attention=True, batchnorm=True)
ma, mb = RM.build(a, "unet", dev, 1e-6, 2.0), RM.build(b, "unet", dev, 1e-6, 2.0)
batch = RM.batch(2, 3, dev)
def run(m, cudnn=True):
    m.train(); m.zero_grad()
    with torch.backends.cudnn.flags(enabled=cudnn):
        _, losses, _ = m.run_on_batch(batch, None, False)
        sum(losses.values()).backward()
    return {k: float(v.detach()) for k, v in losses.items()}, {n: p.grad.detach().double().clone() for n, p in m.named_parameters() if p.grad is not None}
la, ga = run(ma, True)
la2, ga2 = run(ma, True)
ln, gn = run(ma, False)
lo, go = run(mb, True)
def norm(g): return float(torch.sqrt(sum((v ** 2).sum() for v in g.values())))
def dist(g, h): return float(torch.sqrt(sum(((g[k] - h[k]) ** 2).sum() for k in g)))
print("losses cudnn ", la); print("losses native", ln); print("losses rvb   ", lo)
print("|g| cudnn %.6f  native %.6f  rvb %.6f" % (norm(ga), norm(gn), norm(go)))
print("|g_cudnn - g_cudnn(rerun)| / |g| = %.3e" % (dist(ga, ga2) / norm(ga)))
print("|g_cudnn - g_native| / |g| = %.3e" % (dist(ga, gn) / norm(ga)))
print("|g_cudnn - g_rvb|    / |g| = %.3e" % (dist(ga, go) / norm(ga)))
print("|g_native - g_rvb|   / |g| = %.3e" % (dist(gn, go) / norm(ga)))
worst = sorted(((float((ga[k] - go[k]).norm() / (ga[k].norm() + 1e-30)), float((ga[k] - gn[k]).norm() / (ga[k].norm() + 1e-30)), k) for k in ga), reverse=True)[:8]
for r, rn, k in worst: print("  %-60s rvb %.3e  native %.3e" % (k, r, rn))
