"""Bring-up helper for the tcgen05 pixel<->pixel kernel: compares slcl_p2p_fwd / slcl_p2p_bwd with a
plain fp32 torch evaluation on the same bf16-rounded rows (GPU).  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import torch
from slcl import ops  # noqa
op = torch.ops.slcl
dev = torch.device("cuda:0")

def ref(a, b, la, lb, ia, ib, w, T):
    a = a.float().clone().requires_grad_(True); b = b.float().clone().requires_grad_(True)
    s = a @ b.t() / T
    notself = (ia.view(-1, 1) != ib.view(1, -1)).float()
    pos = (la.view(-1, 1) == lb.view(1, -1)).float() * notself
    e = torch.exp(s) * notself
    lp = s - torch.log(e.sum(1, keepdim=True))
    row = -(pos * lp).sum(1) / pos.sum(1)
    loss = (row * w).sum()
    loss.backward()
    return loss.detach(), a.grad, b.grad

def run(A, M, d, T, seed=0, same=False):
    g = torch.Generator(device=dev).manual_seed(seed)
    dp = (d + 63) // 64 * 64
    b = torch.nn.functional.normalize(torch.randn(M, d, device=dev, generator=g), dim=1)
    bb = torch.zeros(M, dp, device=dev, dtype=torch.bfloat16); bb[:, :d] = b.to(torch.bfloat16)
    lb = torch.randint(0, 4, (M,), device=dev, generator=g, dtype=torch.int32)
    ib = torch.arange(M, device=dev, dtype=torch.int32)
    if same:
        ab, la, ia = bb, lb, ib
    else:
        pick = torch.randperm(M, device=dev, generator=g)[:A]
        ab, la, ia = bb[pick].contiguous(), lb[pick].contiguous(), ib[pick].contiguous()
    A = ab.shape[0]
    fg = (la != 0).float(); w = fg / fg.sum()
    shift = torch.full((A,), 1.0 / T, device=dev)
    ma = torch.stack([la, ia], 1).contiguous(); mb = torch.stack([lb, ib], 1).contiguous()
    loss, stats = op.p2p_fwd(ab, bb, ma, mb, shift, w, T)
    torch.cuda.synchronize()
    l_ref, da_ref, db_ref = ref(ab[:, :d], bb[:, :d], la, lb, ia, ib, w, T)
    print(f"A={A} M={M} d={d} T={T} same={same}: loss {loss.item():.6f} ref {l_ref.item():.6f} rel {abs(loss.item()-l_ref.item())/abs(l_ref.item()):.2e}", flush=True)
    da, db = op.p2p_bwd(ab, bb, d, ma, mb, shift, w, T, stats, torch.ones(1, device=dev), True, True)
    torch.cuda.synchronize()
    ea = (da - da_ref).abs().max().item() / da_ref.abs().max().item()
    eb = (db - db_ref).abs().max().item() / db_ref.abs().max().item()
    print(f"   dA rel-to-max err {ea:.2e}   dB rel-to-max err {eb:.2e}", flush=True)

if __name__ == "__main__":
    run(128, 64, 64, 0.7)
    run(128, 128, 64, 0.7)
    run(100, 300, 32, 0.7)
    run(256, 1000, 128, 0.1)
    run(0, 500, 256, 0.07, same=True)
    run(4096, 16384, 256, 0.7)
