"""Kernels that changed after profiles/r2_ncu_full_summary.csv was taken (second ncu pass of round 2): the fused target
step with 64-pixel stages (cfg2) and one BlockConLoss step at (1, 2, 32, 224, 224) -- batched analytic sweeps, table
kernels, the d/8-lanes-per-row finishing kernels, thread-per-row gather / scatter, the exp-shift kernel.

    ncu --set full --import-source on --clock-control none \
        -k regex:'target_tile|p2p_|gather_rows|scatter_rows' -o /tmp/r2b python tools/r2b_profile_targets.py
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import slcl.ops  # noqa
from slcl.loss import BlockConLoss
op = torch.ops.slcl
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
reps = int(os.environ.get("SLCL_PROFILE_REPS", "2"))
b, c, h, k = 32, 128, 256, 5
f = torch.randn(b, c, h, h, device=dev, generator=g)
cen = torch.randn(k, c, device=dev, generator=g)
for _ in range(reps):
    op.target_step(f, cen, 0.25, False, k, 0.1, 1.0, 0.2, False, None, 0.9)
torch.cuda.synchronize()
del f
fb = torch.nn.functional.normalize(torch.randn(1, 2, 32, 224, 224, device=dev, generator=g), dim=2).requires_grad_(True)
lbk = torch.randint(0, 4, (1, 2, 224, 224), device=dev, generator=g)
crit = BlockConLoss(0.7, 32)
for _ in range(reps):
    crit(fb, lbk).backward()
    fb.grad = None
torch.cuda.synchronize()
print("ok")
