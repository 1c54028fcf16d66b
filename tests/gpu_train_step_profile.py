"""Diagnostic (GPU): where the reference UNet's training iteration spends its time behind install(attention=True)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_loader as RL
from reconvat_b200 import synth
dev = torch.device("cuda:0")
B, frames = 8, 640
L = frames * 512
def batch(seed):
    audio = np.stack([(synth.music_int16 if (b & 1) else synth.white_int16)(L, seed * 100 + b) for b in range(B)])
    g = torch.Generator().manual_seed(seed)
    return {"audio": torch.from_numpy(synth.to_float(audio)).to(dev),
            "onset": (torch.rand(B, frames, 88, generator=g) > 0.99).float().to(dev),
            "frame": (torch.rand(B, frames, 88, generator=g) > 0.95).float().to(dev)}
ns = RL.load_patched(attention=True, batchnorm=True)
torch.manual_seed(0)
model = ns.self_attention_VAT.UNet((2, 2), (2, 2), log=True, reconstruction=True, mode="imagewise", spec="Mel", XI=1e-6, eps=2).to(dev).train()
opt = torch.optim.Adam(model.parameters(), 1e-3)
bl, bu = batch(1), batch(2)
def one():
    opt.zero_grad()
    _, losses, _ = model.run_on_batch(bl, bu, True)
    loss = sum(v / 2 if k.startswith("loss/train_LDS") else v for k, v in losses.items())
    loss.backward(); opt.step()
for _ in range(3): one()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): one()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
