"""Host-side cost of one custom-op call through torch.ops.slcl vs the eager dispatch shortcut (slcl.ops.dispatch).
"""
import sys, time, torch
sys.path.insert(0, 'soft-labeled-contrastive-learning_b200')
from slcl import ops, functional as SF
from slcl.loss import ContrastiveLoss
dev = torch.device('cuda:0')
f = torch.randn(2, 32, 16, 16, device=dev); lab = torch.randint(0, 4, (2*16*16,), device=dev)
cs = torch.randn(4, 32, device=dev, requires_grad=True); ct = torch.randn(4, 32, device=dev, requires_grad=True)
def t(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("torch.ops.slcl.class_sums      %.1f us" % t(lambda: torch.ops.slcl.class_sums(f, lab, None, False, 0.0, None, 1, 4)))
print("ops.class_sums (CustomOpDef)   %.1f us" % t(lambda: ops.class_sums(f, lab, None, False, 0.0, None, 1, 4)))
print("ops.class_sums._init_fn        %.1f us" % t(lambda: ops.class_sums._init_fn(f, lab, None, False, 0.0, None, 1, 4)))
crit = ContrastiveLoss()
print("ContrastiveLoss fwd            %.1f us" % t(lambda: crit(cs, ct)))
def fb():
    l = crit(cs, ct); l.backward(); cs.grad = None; ct.grad = None
print("ContrastiveLoss fwd+bwd        %.1f us" % t(fb))
print("empty torch add                %.1f us" % t(lambda: cs + ct))
