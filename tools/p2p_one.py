"""A few forward/backward launches of the analytic p2p path on one shape (the command ncu captures):
   python tools/p2p_one.py A M d            e.g.  python tools/p2p_one.py 4096 16384 256"""
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
sc, sr = ops.self_maps(ia, ib)
fg = (la != 0).float(); w = fg / fg.sum()
shift = torch.full((A,), 1.0 / T, device=dev)
one = torch.ones(1, device=dev)
for _ in range(3):
    loss, stats, state = op.p2p_fwd(ab, bb, ma, mb, shift, w, T, 5, sc, True)
    op.p2p_bwd(ab, bb, d, ma, mb, shift, w, T, stats, one, True, True, 5, sc, sr, state)
torch.cuda.synchronize()
print("ok", float(loss))
