"""Step-by-step smoke of the class-sum kernel modes (each printed before it is launched, flushed)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import slcl.ops  # noqa
op = torch.ops.slcl
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
for (b, c, h, w, k) in ((2, 32, 16, 16, 4), (2, 32, 24, 20, 4), (4, 128, 64, 64, 5), (3, 20, 12, 12, 5)):
    n = b * h * w
    f = torch.randn(b, c, h, w, device=dev, generator=g)
    lab = torch.randint(0, k, (n,), device=dev, generator=g)
    pr = torch.softmax(3 * torch.randn(b, k, h, w, device=dev, generator=g), 1)
    for parts in (1, 2):
        part = (torch.randperm(n, device=dev, generator=g) % parts).to(torch.int32) if parts > 1 else None
        for name, fn in (("hard", lambda: op.class_sums(f, lab, None, False, 0.0, part, parts, k)),
                         ("soft", lambda: op.class_sums(f, None, pr, True, 0.5, part, parts, k)),
                         ("argmax", lambda: op.class_sums(f, None, pr, False, 0.0, part, parts, k))):
            print(f"shape {(b, c, h, w, k)} P={parts} {name} ...", end="", flush=True)
            s = fn()
            torch.cuda.synchronize()
            # fp64 torch reference
            wts = torch.zeros(n, parts * k, dtype=torch.float64, device=dev)
            pid = part.long() if part is not None else torch.zeros(n, dtype=torch.long, device=dev)
            if name == "hard":
                wts[torch.arange(n, device=dev), pid * k + lab] = 1
            else:
                p2 = pr.permute(0, 2, 3, 1).reshape(n, k).double()
                cert = (p2.max(1).values >= 0.5).double() if name == "soft" else torch.ones(n, dtype=torch.float64, device=dev)
                src = p2 if name == "soft" else torch.nn.functional.one_hot(p2.argmax(1), k).double()
                for kk in range(k):
                    wts[torch.arange(n, device=dev), pid * k + kk] = src[:, kk] * cert
            x = f.permute(0, 2, 3, 1).reshape(n, c).double()
            ref = torch.cat([wts.t() @ x, wts.sum(0, keepdim=True).t()], 1)
            err = float((s - ref).abs().max() / ref.abs().max())
            print(f" ok, rel err {err:.2e}", flush=True)
            assert err < 1e-5
print("all ok")
