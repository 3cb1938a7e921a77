"""f-1 source side at the cfg4 per-GPU shape (16 x 32 x 224 x 224, K4; the map is 98 MiB < 126 MB L2):
update_class_center_iter + source mpcl_loss_calc forward + backward as one call (slcl.loss.mpcl_source_step).
Prints the step time as a CUDA graph; under ncu (SLCL_PROBE_ONCE=1) it flushes L2, then runs ONE step so the DRAM bytes
of its kernels show how many of the three walks over F_s were served by L2."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from slcl.loss import MPCL, mpcl_source_step
dev = torch.device("cuda:0")
b, c, h, k = 16, 32, 224, 4
g = torch.Generator(device=dev).manual_seed(1)
f = torch.randn(b, c, h, h, device=dev, generator=g).requires_grad_(True)
lab = torch.randint(0, k, (b, h, h), device=dev, generator=g)
cen = torch.randn(k, c, device=dev, generator=g)
mp = MPCL(dev, num_class=k, temperature=0.1, m=0.4, base_temperature=1.0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def step():
    centres, loss = mpcl_source_step(f, lab, cen, mp, m=0.9, num_class=k)
    loss.backward()
    f.grad = None
    return loss.detach()


if os.environ.get("SLCL_PROBE_ONCE") == "1":
    step(); torch.cuda.synchronize()
    flush.fill_(1); torch.cuda.synchronize()           # L2 now holds the flush buffer, not F_s
    step(); torch.cuda.synchronize()
    print("once ok")
    sys.exit(0)

for _ in range(3):
    step()
side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream(dev).wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    keep = step()
for name, fn in (("eager", step), ("one CUDA graph", graph.replay)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(30):
        fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 30
    n = b * h * h
    print(f"source step {name}: {ms*1e3:.1f} us; algorithmic (4C+8) + (12C+16) B/px = {(16*c+24)} B/px -> {(16*c+24)*n/ms/1e6:.0f} GB/s")
