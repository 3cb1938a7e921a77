"""Which pieces of the MCCL loss section survive CUDA-graph capture."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from slcl.loss import ContrastiveLoss, cnr_loss
from slcl.utils_ import cal_centroid
dev = torch.device("cuda:0")
b, c, h, k = 4, 32, 32, 4
gen = torch.Generator(device=dev).manual_seed(1)
ft = torch.randn(b, c, h, h, device=dev, generator=gen).requires_grad_(True)
lab = torch.randint(0, k, (b, h, h), device=dev, generator=gen)
pr = torch.softmax(torch.randn(b, k, h, h, device=dev, generator=gen), 1).requires_grad_(True)
part = (torch.randperm(b * h * h, device=dev, generator=gen) % 2).to(torch.int32)
crit = ContrastiveLoss()
cs0 = torch.randn(k, c, device=dev, requires_grad=True); ct0 = torch.randn(k, c, device=dev, requires_grad=True)
def hard_fwd(): return cal_centroid(ft, lab, n_class=k)[0].sum()
def hard_fb():
    l = cal_centroid(ft, lab, n_class=k)[0].sum(); l.backward(); ft.grad = None; return l
def soft_fb():
    ct = cal_centroid(ft, pr, pseudo_label=True, weighted_ave=True, partition=2, n_class=k, part_id=part)[0]
    l = ct[0].sum() + ct[1].sum(); l.backward(); ft.grad = None; pr.grad = None; return l
def crit_fb():
    l = crit(cs0, ct0) + 4e-5 * cnr_loss(cs0, [ct0]); l.backward(); cs0.grad = None; ct0.grad = None; return l
for name, fn in (("hard fwd", hard_fwd), ("hard fwd+bwd", hard_fb), ("soft fwd+bwd", soft_fb), ("crit fwd+bwd", crit_fb)):
    try:
        side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "OK", float(out))
    except Exception as e:
        print(name, "FAILED", str(e).splitlines()[0])
        torch.cuda.synchronize()
