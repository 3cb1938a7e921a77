"""One launch (after a warm-up) of every kernel whose ncu --set full row is committed under profiles/ for round 2:
prototype forward / backward (cfg2), class sums hard + soft x 2 partitions and centroid backward (cfg5 geometry), the
one-pass fused target step (cfg2), segmentation losses, entropy map, sampler (compaction, gather), pixel<->pixel sweeps
(analytic and general, cfg3).

    ncu --set full --import-source on --clock-control none -k regex:'proto_|class_sums_v3|centroid_bwd4|target_tile|seg_|entropy|compact|gather_rows|p2p_kernel' \
        -o gpurun_out/r2_full python tools/r2_profile_targets.py
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import slcl.ops as slcl_ops  # noqa
from slcl.plan import P2PPlan, ProtoPlan
from slcl import seg
op = torch.ops.slcl
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
reps = int(os.environ.get("SLCL_PROFILE_REPS", "2"))


def run(fn):
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return out


# cfg2: prototype loss forward / backward, fused target step
b, c, h, k = 32, 128, 256, 5
n = b * h * h
f = torch.randn(b, c, h, h, device=dev, generator=g)
lab = torch.randint(0, k, (n,), device=dev, generator=g)
sel = (torch.rand(n, device=dev, generator=g) > 0.5).float()
cen = torch.randn(k, c, device=dev, generator=g)
plan = ProtoPlan(f, lab, sel, cen, k, 0.1, 1.0, 0.2)
run(lambda: (plan.forward(), plan.backward()))
run(lambda: op.target_step(f, cen, 0.25, False, k, 0.1, 1.0, 0.2, False, None, 0.9))
run(lambda: op.pseudo_label(f, cen, 0.25))
del f, lab, sel, plan
# cfg5 geometry: class sums and centroid backward
b, c, h, k, parts = 64, 32, 224, 4, 2
n = b * h * h
f = torch.randn(b, c, h, h, device=dev, generator=g)
lab = torch.randint(0, k, (n,), device=dev, generator=g)
pr = torch.softmax(3 * torch.randn(b, k, h, h, device=dev, generator=g), 1)
part = (torch.randperm(n, device=dev, generator=g) % parts).to(torch.int32)
gc = torch.randn(parts * k, c, device=dev, generator=g)
run(lambda: op.class_sums(f, lab, None, False, 0.0, None, 1, k))
sums = run(lambda: op.class_sums(f, None, pr, True, 0.0, part, parts, k))
run(lambda: op.centroid_bwd(f, None, pr, True, 0.0, part, parts, k, gc, sums, 1.0, True))
# segmentation losses / entropy map on [64, 4, 224, 224] logits
z = torch.randn(b, k, h, h, device=dev, generator=g).requires_grad_(True)
lz = lab.view(b, h, h)


def seg_step():
    z.grad = None
    out = seg.loss_calc(z, lz, 0, True) + seg.dice_loss(z, lz)
    out.backward()
run(seg_step)
run(lambda: op.entropy_map(pr))
# sampler
run(lambda: op.compact_by_class(lab, k))
fm = torch.randn(16, 256, 64, 64, device=dev, generator=g)
rows = torch.randperm(16 * 64 * 64, device=dev, generator=g)[:20480]
run(lambda: op.gather_unit_rows(fm, rows, True, True, False))
del f, pr, part, z, fm
# cfg3 pixel <-> pixel: analytic and general sweeps
A, M, d, T = 4096, 16384, 256, 0.7
bb = torch.nn.functional.normalize(torch.randn(M, d, device=dev, generator=g), dim=1).to(torch.bfloat16)
lb = torch.randint(0, 5, (M,), device=dev, generator=g, dtype=torch.int32)
ib = torch.arange(M, device=dev, dtype=torch.int32)
pick = torch.randperm(M, device=dev, generator=g)[:A]
a, la, ia = bb[pick].contiguous(), lb[pick].contiguous(), ib[pick].contiguous()
fg = (la != 0).float()
shift, weight = torch.full((A,), 1.0 / T, device=dev), fg / fg.sum()
ma, mb = slcl_ops.pad_meta(la, ia), slcl_ops.pad_meta(lb, ib)
sc, sr = slcl_ops.self_maps(ia, ib)
pa = P2PPlan(a, bb, d, ma, mb, shift, weight, T, n_class=5, a_selfcol=sc, b_selfrow=sr)
pg = P2PPlan(a, bb, d, ma, mb, shift, weight, T)
run(lambda: (pa.forward(), pa.backward()))
run(lambda: (pg.forward(), pg.backward()))
print("ok")
