"""Host-side cost of the pixel<->pixel Python path (BlockConLoss at the reference shape): op counts and CPU time."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from slcl.loss import BlockConLoss, SupConLoss
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
fb = torch.nn.functional.normalize(torch.randn(1, 2, 32, 224, 224, device=dev, generator=g), dim=2).requires_grad_(True)
lbk = torch.randint(0, 4, (1, 2, 224, 224), device=dev, generator=g)
crit = BlockConLoss(0.7, 32)


def step():
    loss = crit(fb, lbk)
    loss.backward()
    fb.grad = None


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    step()
t_host = (time.perf_counter() - t0) / 50
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"host time per step (launch side only) {t_host*1e6:.0f} us")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=22, max_name_column_width=90))
