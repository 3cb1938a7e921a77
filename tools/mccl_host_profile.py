"""Host-side (Python) cost of the MCCL loss section in eager mode: cProfile over 100 steps."""
import cProfile, pstats, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from slcl.loss import ContrastiveLoss, cnr_loss
from slcl.utils_ import cal_centroid
dev = torch.device("cuda:0")
b, c, h, k, parts = 8, 32, 64, 4, 2          # small maps: the device side is negligible, the host side is what is timed
n_px = b * h * h
gen = torch.Generator(device=dev).manual_seed(1)
ft = [torch.randn(b, c, h, h, device=dev, generator=gen).requires_grad_(True) for _ in range(3)]
lab_s = torch.randint(0, k, (b, h, h), device=dev, generator=gen)
pr = [torch.softmax(3 * torch.randn(b, k, h, h, device=dev, generator=gen), 1).requires_grad_(True) for _ in range(2)]
part = [(torch.randperm(n_px, device=dev, generator=gen) % parts).to(torch.int32) for _ in range(2)]
crit = ContrastiveLoss()


def step():
    cs, _, _ = cal_centroid(ft[0], lab_s, n_class=k)
    loss = 0
    for i in range(2):
        ct, _, _ = cal_centroid(ft[1 + i], pr[i], pseudo_label=True, weighted_ave=True, partition=parts, n_class=k, part_id=part[i])
        for c_p in ct:
            loss = loss + crit(cs, c_p)
        loss = loss + 4e-5 * cnr_loss(cs, ct)
    loss.backward()
    for t in ft + pr:
        t.grad = None


for _ in range(10):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    step()
torch.cuda.synchronize()
print(f"eager step (host-bound at this size): {(time.perf_counter() - t0) * 1e4:.1f} us")
pr_ = cProfile.Profile()
pr_.enable()
for _ in range(100):
    step()
pr_.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr_)
st.sort_stats("cumulative").print_stats(45)
