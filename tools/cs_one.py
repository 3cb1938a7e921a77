"""One launch of the soft P=2 class sums (and the centroid backward) at the cfg5 geometry -- the ncu target."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import slcl.ops  # noqa
op = torch.ops.slcl
dev = torch.device("cuda:0")
b, c, h, k, parts = 64, 32, 224, 4, 2
if len(sys.argv) > 1 and sys.argv[1] == "cfg2":
    b, c, h, k = 32, 128, 256, 5
g = torch.Generator(device=dev).manual_seed(1)
n = b * h * h
f = torch.randn(b, c, h, h, device=dev, generator=g)
pr = torch.softmax(3 * torch.randn(b, k, h, h, device=dev, generator=g), 1)
part = (torch.randperm(n, device=dev, generator=g) % parts).to(torch.int32)
gc = torch.randn(parts * k, c, device=dev, generator=g)
for _ in range(3):
    sums = op.class_sums(f, None, pr, True, 0.5, part, parts, k)
    op.centroid_bwd(f, None, pr, True, 0.5, part, parts, k, gc, sums, 1.0, True)
torch.cuda.synchronize()
print("ok", float(sums.sum()))
