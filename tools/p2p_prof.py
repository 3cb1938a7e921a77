"""bring-up (needs a library built with SLCL_EXTRA_NVCC_FLAGS=-DSLCL_P2P_PROFILE): per-role wait-cycle profile of one p2p sweep (CTA 0)."""
import os, sys, torch
sys.path.insert(0, 'soft-labeled-contrastive-learning_b200')
dev = torch.device('cuda:0')
prof = torch.zeros(16, dtype=torch.int64, device=dev)
os.environ["SLCL_P2P_PROF"] = str(prof.data_ptr())
from slcl import ops
op = torch.ops.slcl
A, M, d, T = 16384, 16384, 256, 0.7
g = torch.Generator(device=dev).manual_seed(1)
bb = torch.nn.functional.normalize(torch.randn(M, d, device=dev, generator=g), dim=1).to(torch.bfloat16)
lb = torch.randint(0, 5, (M,), device=dev, generator=g, dtype=torch.int32)
ib = torch.arange(M, device=dev, dtype=torch.int32)
mb = ops.pad_meta(lb, ib)
w = torch.full((A,), 1.0 / A, device=dev); shift = torch.full((A,), 1.0 / T, device=dev); one = torch.ones(1, device=dev)
names = ["prod total", "prod wait c_empty", "prod wait m_empty", "-", "mma total", "mma wait c_full", "mma wait s_empty", "mma wait g_full",
         "epi4 total", "epi4 wait m_full", "epi4 wait s_full", "epi4 wait g_empty", "epi11 total", "epi11 wait m_full", "epi11 wait s_full", "epi11 wait g_empty"]
for mode in ("fwd", "bwd"):
    for _ in range(2):
        loss, stats, _, _ = op.p2p_fwd(bb, bb, mb, mb, shift, w, T)
        torch.cuda.synchronize()
        if mode == "fwd":
            vals = prof.cpu().tolist()
        else:
            op.p2p_bwd(bb, bb, d, mb, mb, shift, w, T, stats, one, True, False)
            torch.cuda.synchronize()
            vals = prof.cpu().tolist()
    print(mode, "tiles=256; cycles per tile:")
    for n, v in zip(names, vals):
        print(f"   {n:20s} {v/256:9.0f}")
