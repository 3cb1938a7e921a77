"""One shape of the fused target step (for ncu): python tools/target_step_one.py [B C H K]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import slcl.ops  # noqa
op = torch.ops.slcl
dev = torch.device("cuda:0")
b, c, h, k = [int(x) for x in sys.argv[1:5]] if len(sys.argv) >= 5 else (32, 128, 256, 5)
g = torch.Generator(device=dev).manual_seed(1)
f = torch.randn(b, c, h, h, device=dev, generator=g)
cen = torch.randn(k, c, device=dev, generator=g)
for _ in range(3):
    op.target_step(f, cen, 0.25, False, k, 0.1, 1.0, 0.2, False, None, 0.9)
torch.cuda.synchronize()
print("ok")
