"""Fused target step at one shape with SLCL_TILE_TEAMS from the environment (team-count scaling of the pixel warps):
    SLCL_TILE_TEAMS=4 python tools/target_step_teams.py B C H K
"""
import os, sys, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import slcl.ops  # noqa
op = torch.ops.slcl
dev = torch.device("cuda:0")
b, c, h, k = [int(x) for x in sys.argv[1:5]]
g = torch.Generator(device=dev).manual_seed(1)
f = torch.randn(b, c, h, h, device=dev, generator=g)
cen = torch.randn(k, c, device=dev, generator=g)
fn = lambda: op.target_step(f, cen, 0.25, False, k, 0.1, 1.0, 0.2, False, None, 0.9)
for _ in range(3): fn()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): fn()
e.record(); torch.cuda.synchronize()
print(f"teams={os.environ.get('SLCL_TILE_TEAMS')} C={c}: {s.elapsed_time(e)/20*1e3:.1f} us")
