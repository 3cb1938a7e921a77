"""Kernel-level breakdown of the MCCL loss section (same step as tools/mccl_profile.py): device time per kernel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from torch.profiler import profile, ProfilerActivity
from slcl.loss import ContrastiveLoss, cnr_loss
from slcl.utils_ import cal_centroid
dev = torch.device("cuda:0")
b, c, h, k, parts = 64, 32, 224, 4, 2
n_px = b * h * h
gen = torch.Generator(device=dev).manual_seed(1)
ft = [torch.randn(b, c, h, h, device=dev, generator=gen).requires_grad_(True) for _ in range(3)]
lab_s = torch.randint(0, k, (b, h, h), device=dev, generator=gen)
pr = [torch.softmax(3 * torch.randn(b, k, h, h, device=dev, generator=gen), 1).requires_grad_(True) for _ in range(2)]
part = [(torch.randperm(n_px, device=dev, generator=gen) % parts).to(torch.int32) for _ in range(2)]
crit = ContrastiveLoss()


def step():
    cs, _, _ = cal_centroid(ft[0], lab_s, n_class=k)
    loss = 0
    for i in range(2):
        ct, _, _ = cal_centroid(ft[1 + i], pr[i], pseudo_label=True, weighted_ave=True, partition=parts, n_class=k, part_id=part[i])
        for c_p in ct:
            loss = loss + crit(cs, c_p)
        loss = loss + 4e-5 * cnr_loss(cs, ct)
    loss.backward()
    for t in ft + pr:
        t.grad = None


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
rows = sorted(((e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[1])
print(f"device time per step {sum(r[1] for r in rows):.0f} us in {sum(r[2] for r in rows):.0f} kernels")
for kname, t, n in rows[:30]:
    print(f"{t:8.1f} us x{n:4.1f}  {kname[:120]}")
