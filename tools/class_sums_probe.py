"""Which role limits class_sums_v3 at the cfg5 geometry (C=32, K=4)?  Times variants that load the builder and the
consumers differently: hard/soft labels x 1/2 partitions (+ cfg2 geometry)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import slcl.ops  # noqa
op = torch.ops.slcl
dev = torch.device("cuda:0")
PEAK = 6521.4


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (b, c, h, k) in ((64, 32, 224, 4), (32, 128, 256, 5)):
    g = torch.Generator(device=dev).manual_seed(1)
    n = b * h * h
    f = torch.randn(b, c, h, h, device=dev, generator=g)
    lab = torch.randint(0, k, (n,), device=dev, generator=g)
    pr = torch.softmax(3 * torch.randn(b, k, h, h, device=dev, generator=g), 1)
    gcen = torch.randn(2 * k, c, device=dev, generator=g)
    for parts in (1, 2, 3):
        part = (torch.randperm(n, device=dev, generator=g) % parts).to(torch.int32) if parts > 1 else None
        for name, fn, byt in (
                ("hard", lambda: op.class_sums(f, lab, None, False, 0.0, part, parts, k), 4 * c + 8 + (4 if part is not None else 0)),
                ("soft", lambda: op.class_sums(f, None, pr, True, 0.0, part, parts, k), 4 * c + 4 * k + (4 if part is not None else 0)),
                ("soft+thr", lambda: op.class_sums(f, None, pr, True, 0.5, part, parts, k), 4 * c + 4 * k + (4 if part is not None else 0)),
                ("argmax", lambda: op.class_sums(f, None, pr, False, 0.0, part, parts, k), 4 * c + 4 * k + (4 if part is not None else 0))):
            if parts * k > 16:
                continue
            ms = timed(fn)
            print(f"C={c} K={k} P={parts} {name:9s} {ms*1e3:8.1f} us  {byt*n/ms/1e6:7.0f} GB/s  frac {byt*n/ms/1e6/PEAK:.3f}")
    if parts * k <= 16:
        pass
    for parts in (1, 2):
        part = (torch.randperm(n, device=dev, generator=g) % parts).to(torch.int32) if parts > 1 else None
        sums = op.class_sums(f, None, pr, True, 0.0, part, parts, k)
        gc = torch.randn(parts * k, c, device=dev, generator=g)
        for dp in (True, False):
            ms = timed(lambda: op.centroid_bwd(f, None, pr, True, 0.0, part, parts, k, gc, sums, 1.0, dp))
            byt = (8 * c + (8 * k if dp else 4 * k) + (4 if part is not None else 0))
            print(f"C={c} K={k} P={parts} centroid_bwd dP={dp}  {ms*1e3:8.1f} us  {byt*n/ms/1e6:7.0f} GB/s  frac {byt*n/ms/1e6/PEAK:.3f}")
        sums_h = op.class_sums(f, lab, None, False, 0.0, None, 1, k)
        ms = timed(lambda: op.centroid_bwd(f, lab, None, False, 0.0, None, 1, k, gc[:k].contiguous(), sums_h, 1.0, False))
        print(f"C={c} K={k} hard centroid_bwd (write-only dF)  {ms*1e3:8.1f} us  {(4*c+8)*n/ms/1e6:7.0f} GB/s  frac {(4*c+8)*n/ms/1e6/PEAK:.3f}")
