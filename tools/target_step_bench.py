"""Fused target step (one pass) vs the separate calls, cfg2 / cfg4 / cfg5 shapes."""
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


for (b, c, h, k) in ((32, 128, 256, 5), (16, 32, 224, 4), (64, 32, 224, 4), (32, 64, 256, 5)):
    g = torch.Generator(device=dev).manual_seed(1)
    n = b * h * h
    f = torch.randn(b, c, h, h, device=dev, generator=g)
    cen = torch.randn(k, c, device=dev, generator=g)
    ms_f = timed(lambda: op.target_step(f, cen, 0.25, False, k, 0.1, 1.0, 0.2, False, None, 0.9))

    def sep():
        out = op.proto_fwd_target(f, cen, 0.25, k, 0.1, 1.0, 0.2, False)
        op.centroids_fwd(f, out[3], None, False, 0.0, None, 1, k, None, 0.9)
    ms_s = timed(sep)
    byt = (4 * c) * n            # ONE read of the map (+ 12 B/px label/sel out, 4(K+1) stash)
    print(f"B={b} C={c} {h}x{h} K={k}: fused {ms_f*1e3:.1f} us ({byt/ms_f/1e6:.0f} GB/s of map bytes = {byt/ms_f/1e6/PEAK:.3f} of peak)   "
          f"separate {ms_s*1e3:.1f} us   speed-up {ms_s/ms_f:.2f}x", flush=True)
