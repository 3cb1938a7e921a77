"""Kernel-level breakdown of the cfg3 public-API step (sampled_supcon_loss fwd+bwd): torch profiler table."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from slcl.p2p import sampled_supcon_loss
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
fmap = torch.randn(16, 256, 64, 64, device=dev, generator=g).requires_grad_(True)
lmap = torch.randint(0, 5, (16, 64, 64), device=dev, generator=g)


def step():
    loss = sampled_supcon_loss(fmap, lmap, 4096, 16384, 5, temperature=0.7)
    loss.backward()
    fmap.grad = None


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
rows = sorted(((e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"device time per step {tot:.0f} us in {sum(r[2] for r in rows):.0f} kernels")
for k, t, n in rows[:45]:
    print(f"{t:8.1f} us x{n:4.1f}  {k[:130]}")
