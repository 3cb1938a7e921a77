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
    ma = ops.pad_meta(la, ia); mb = ops.pad_meta(lb, ib)
    loss, stats, _ = op.p2p_fwd(ab, bb, ma, mb, shift, w, T)
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


def bench(A=4096, M=16384, d=256, T=0.7, iters=20):
    g = torch.Generator(device=dev).manual_seed(1)
    bb = torch.nn.functional.normalize(torch.randn(M, d, device=dev, generator=g), dim=1).to(torch.bfloat16)
    lb = torch.randint(0, 5, (M,), device=dev, generator=g, dtype=torch.int32)
    ib = torch.arange(M, device=dev, dtype=torch.int32)
    pick = torch.randperm(M, device=dev, generator=g)[:A]
    ab, la, ia = bb[pick].contiguous(), lb[pick].contiguous(), ib[pick].contiguous()
    fg = (la != 0).float(); w = fg / fg.sum()
    shift = torch.full((A,), 1.0 / T, device=dev)
    ma = ops.pad_meta(la, ia); mb = ops.pad_meta(lb, ib)
    one = torch.ones(1, device=dev)
    def fwd(): return op.p2p_fwd(ab, bb, ma, mb, shift, w, T)
    loss, stats = fwd()
    def bwd(): return op.p2p_bwd(ab, bb, d, ma, mb, shift, w, T, stats, one, True, True)
    def timed(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters): fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / iters * 1e3
    tf, tb = timed(fwd), timed(bwd)
    fl = 2.0 * A * M * d
    print(f"p2p A={A} M={M} d={d}: fwd {tf:.1f} us ({fl/tf/1e6:.0f} TFLOP/s)  bwd {tb:.1f} us ({3*fl/tb/1e6:.0f} TFLOP/s of 6AMd)  "
          f"total {tf+tb:.1f} us -> {4*fl/(tf+tb)/1e6:.0f} TFLOP/s algorithmic (8AMd), {4*fl/(tf+tb)/1e6/1643.3*100:.1f}% of 1643 TF peak")

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "bench":
    bench()
    bench(16384, 16384, 256)
    bench(8192, 65536, 128)
