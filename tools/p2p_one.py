"""bring-up: a few forward/backward launches of the p2p op on one shape (for ncu)."""
import sys, torch
sys.path.insert(0, 'soft-labeled-contrastive-learning_b200')
from slcl import ops
op = torch.ops.slcl
dev = torch.device('cuda:0')
A, M, d, T = [int(x) for x in sys.argv[1:4]] + [0.7]
g = torch.Generator(device=dev).manual_seed(1)
bb = torch.nn.functional.normalize(torch.randn(M, d, device=dev, generator=g), dim=1).to(torch.bfloat16)
lb = torch.randint(0, 5, (M,), device=dev, generator=g, dtype=torch.int32)
ib = torch.arange(M, device=dev, dtype=torch.int32)
pick = torch.randperm(M, device=dev, generator=g)[:A]
ab, la, ia = bb[pick].contiguous(), lb[pick].contiguous(), ib[pick].contiguous()
ma, mb = ops.pad_meta(la, ia), ops.pad_meta(lb, ib)
fg = (la != 0).float(); w = fg / fg.sum()
shift = torch.full((A,), 1.0 / T, device=dev)
one = torch.ones(1, device=dev)
for _ in range(3):
    loss, stats, _ = op.p2p_fwd(ab, bb, ma, mb, shift, w, T)
    if len(sys.argv) > 4:
        op.p2p_bwd(ab, bb, d, ma, mb, shift, w, T, stats, one, True, True)
torch.cuda.synchronize()
print("ok", float(loss))
