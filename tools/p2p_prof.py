"""bring-up: per-role wait-cycle profile of the analytic p2p sweeps on cfg3 (CTA 0 of each sweep).
Needs a library built with the in-kernel counters:
   SLCL_LIB_NAME=libslcl_prof.so SLCL_EXTRA_NVCC_FLAGS=-DSLCL_P2P_PROFILE python soft-labeled-contrastive-learning_b200/build.py
   SLCL_LIB_PATH=$PWD/soft-labeled-contrastive-learning_b200/slcl/libslcl_prof.so python tools/p2p_prof.py"""
import os, sys, torch
sys.path.insert(0, 'soft-labeled-contrastive-learning_b200')
dev = torch.device('cuda:0')
prof = torch.zeros(16, dtype=torch.int64, device=dev)
os.environ["SLCL_P2P_PROF"] = str(prof.data_ptr())
from slcl import ops
op = torch.ops.slcl
A, M, d, T = 4096, 16384, 256, 0.7
g = torch.Generator(device=dev).manual_seed(1)
bb = torch.nn.functional.normalize(torch.randn(M, d, device=dev, generator=g), dim=1).to(torch.bfloat16)
lb = torch.randint(0, 5, (M,), device=dev, generator=g, dtype=torch.int32)
ib = torch.arange(M, device=dev, dtype=torch.int32)
pick = torch.randperm(M, device=dev, generator=g)[:A]
ab, la, ia = bb[pick].contiguous(), lb[pick].contiguous(), ib[pick].contiguous()
ma, mb = ops.pad_meta(la, ia), ops.pad_meta(lb, ib)
sc, sr = ops.self_maps(ia, ib)
w = torch.full((A,), 1.0 / A, device=dev); shift = torch.full((A,), 1.0 / T, device=dev); one = torch.ones(1, device=dev)
names = ["prod total", "prod wait c_empty", "prod wait m_empty", "SETUP+R (abs)", "mma total", "mma wait c_full", "mma wait s_empty", "mma wait g_full",
         "epi4 total", "  of SETUP: init+alloc (abs)", "  of SETUP: pdl wait (abs)", "DRAIN (abs)", "epi11 total", "  of DRAIN: acc_full wait (abs)", "epi11 wait s_full", "CTA TOTAL (abs)"]
tiles = {"fwd only": 64, "fwd+U": 64, "dB": 64}
for mode in ("fwd only", "fwd+U", "dB"):
    for _ in range(2):
        prof.zero_()
        loss, stats, state = op.p2p_fwd(ab, bb, ma, mb, shift, w, T, 5, sc, mode != "fwd only")
        torch.cuda.synchronize()
        if mode == "dB":
            prof.zero_()
            op.p2p_bwd(ab, bb, d, ma, mb, shift, w, T, stats, one, True, True, 5, sc, sr, state)
            torch.cuda.synchronize()
        vals = prof.cpu().tolist()
    print(mode, f"cycles per tile ({tiles[mode]} tiles per CTA):")
    for n, v in zip(names, vals):
        if "abs" in n:
            print(f"   {n:20s} {v:9.0f} cycles")
        elif n != "-":
            print(f"   {n:20s} {v / tiles[mode]:9.0f}")
