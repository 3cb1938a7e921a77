"""bring-up (needs a library built with SLCL_EXTRA_NVCC_FLAGS=-DSLCL_P2P_PROFILE): forward-sweep wait-cycle profile under a debug knob."""
import os, sys, torch
sys.path.insert(0, 'soft-labeled-contrastive-learning_b200')
dev = torch.device('cuda:0')
prof = torch.zeros(16, dtype=torch.int64, device=dev)
os.environ["SLCL_P2P_PROF"] = str(prof.data_ptr())
from slcl import ops
op = torch.ops.slcl
A = M = 16384; d = 256; T = 0.7
g = torch.Generator(device=dev).manual_seed(1)
bb = torch.nn.functional.normalize(torch.randn(M, d, device=dev, generator=g), dim=1).to(torch.bfloat16)
lb = torch.randint(0, 5, (M,), device=dev, generator=g, dtype=torch.int32)
mb = ops.pad_meta(lb, torch.arange(M, device=dev, dtype=torch.int32))
w = torch.full((A,), 1.0 / A, device=dev); shift = torch.full((A,), 1.0 / T, device=dev)
for _ in range(2):
    op.p2p_fwd(bb, bb, mb, mb, shift, w, T); torch.cuda.synchronize()
v = prof.cpu().tolist()
print("debug", os.environ.get("SLCL_P2P_DEBUG", "0"), "per tile: mma total %d (wait s_empty %d) epi total %d (wait s_full %d)" % (v[4]/256, v[6]/256, v[8]/256, v[10]/256))
