"""Prototype loss fwd+bwd at the cfg4 per-GPU shape(s): step time, and the per-kernel split, with/without the L2 hint."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from slcl.plan import ProtoPlan
dev = torch.device("cuda:0")
PEAK = 6521.4


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (b, c, h, k) in ((16, 32, 224, 4), (32, 32, 224, 4), (64, 32, 224, 4), (128, 32, 224, 4), (2, 128, 256, 5)):
    g = torch.Generator(device=dev).manual_seed(1)
    n = b * h * h
    f = torch.randn(b, c, h, h, device=dev, generator=g)
    lab = torch.randint(0, k, (n,), device=dev, generator=g)
    sel = (torch.rand(n, device=dev, generator=g) > 0.3).float()
    cen = torch.randn(k, c, device=dev, generator=g)
    plan = ProtoPlan(f, lab, sel, cen, k, 0.1, 1.0, 0.2)

    def step():
        plan.forward(); plan.backward()
    ms = timed(step)
    ms_f = timed(plan.forward)
    ms_b = timed(plan.backward)
    graph = plan.capture_graph()
    ms_g = timed(graph.replay)
    byt = (12 * c + 24) * n
    print(f"B={b} C={c} {h}x{h} K={k} map {n*c*4/2**20:.0f} MiB: step {ms*1e3:.1f} us (graph {ms_g*1e3:.1f})  fwd {ms_f*1e3:.1f}  bwd {ms_b*1e3:.1f}  "
          f"frac {byt/ms/1e6/PEAK:.3f} (graph {byt/ms_g/1e6/PEAK:.3f})", flush=True)
