"""Third ncu pass of round 2: one cfg3 step through the public API (sampled_supcon_loss forward + backward) -- the sampler's
bookkeeping kernels, metadata / weights, lookup-table self maps and the pixel-side backward scatter.

    ncu --set full --import-source on --clock-control none \
        -k regex:'balanced_|gather_labels|compact_|self_|rows_meta|tile_|scatter_by_map|gather_rows' -o /tmp/r2c \
        python tools/r2c_profile_targets.py
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from slcl.p2p import sampled_supcon_loss
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
fmap = torch.randn(16, 256, 64, 64, device=dev, generator=g).requires_grad_(True)
lmap = torch.randint(0, 5, (16, 64, 64), device=dev, generator=g)
for _ in range(int(os.environ.get("SLCL_PROFILE_REPS", "2"))):
    sampled_supcon_loss(fmap, lmap, 4096, 16384, 5, temperature=0.7).backward()
    fmap.grad = None
torch.cuda.synchronize()
print("ok")
