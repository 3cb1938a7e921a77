"""Kernel-tuning helper (not part of the product): time proto fwd / bwd / pseudo-label of one
libslcl variant on the cfg2 shape through the raw C ABI.  Usage on the GPU box:
    SLCL_LIB_PATH=.../libslcl_x.so python tools/microbench_proto.py [B C H W K]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import torch
from slcl.plan import ProtoPlan
from slcl import ops  # noqa

B, C, H, W, K = [int(x) for x in sys.argv[1:6]] if len(sys.argv) >= 6 else (32, 128, 256, 256, 5)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.randn(B, C, H, W, device=dev, generator=g)
lab = torch.randint(0, K, (B * H * W,), device=dev, generator=g)
sel = (torch.rand(B * H * W, device=dev, generator=g) > 0.5).float()
cen = torch.randn(K, C, device=dev, generator=g)
plan = ProtoPlan(feat, lab, sel, cen, K, 0.1, 1.0, 0.2)

def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]

n = B * H * W
f_med, f_min = timed(plan.forward)
b_med, b_min = timed(plan.backward)
def both(): plan.forward(); plan.backward()
s_med, s_min = timed(both)
gb = lambda byt, ms: byt / ms / 1e6
print(f"{os.path.basename(os.environ.get('SLCL_LIB_PATH', 'libslcl.so')):24s} "
      f"fwd {f_med*1e3:7.1f} us ({gb((4*C+12)*n, f_med):6.0f} GB/s)  bwd {b_med*1e3:7.1f} us ({gb(8*C*n, b_med):6.0f} GB/s)  "
      f"step {s_med*1e3:7.1f} us ({gb((12*C+24)*n, s_med):6.0f} GB/s, {n/s_med/1e6:6.2f} Gpix/s)  min f/b/s {f_min*1e3:.1f}/{b_min*1e3:.1f}/{s_min*1e3:.1f}")
