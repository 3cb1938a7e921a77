# bring-up only: forward-sweep bottleneck bisect (SLCL_P2P_DEBUG knobs); every run under a short timeout
for d in 31 63; do echo "debug=$d"; SLCL_P2P_DEBUG=$d timeout 60 python - <<'PY' 2>&1 | grep fwd
import sys, torch
sys.path.insert(0,'tools'); sys.path.insert(0,'soft-labeled-contrastive-learning_b200')
from slcl import ops
op = torch.ops.slcl
dev = torch.device('cuda:0')
A = M = 16384; d = 256; T = 0.7
g = torch.Generator(device=dev).manual_seed(1)
bb = torch.nn.functional.normalize(torch.randn(M, d, device=dev, generator=g), dim=1).to(torch.bfloat16)
lb = torch.randint(0, 5, (M,), device=dev, generator=g, dtype=torch.int32)
ib = torch.arange(M, device=dev, dtype=torch.int32)
mb = ops.pad_meta(lb, ib)
w = torch.full((A,), 1.0 / A, device=dev); shift = torch.full((A,), 1.0 / T, device=dev)
def fwd(): return op.p2p_fwd(bb, bb, mb, mb, shift, w, T)
for _ in range(3): fwd()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): fwd()
e.record(); torch.cuda.synchronize()
print(f"fwd {s.elapsed_time(e)/10*1e3:.1f} us")
PY
done
