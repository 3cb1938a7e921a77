"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).

The CUDA path (called through the drop-in Python API -> torch custom ops ->
C ABI of libslcl.so) is compared with

  * the committed outputs of the reference's own functions
    (tests/golden/reference_outputs.npz), and
  * the CPU oracle (oracle/slcl_oracle.py) on the same seeded inputs.

Tolerances (BASELINE.json north_star): loss and gradients rtol 1e-4 in fp32;
labels, selection masks, class counts and compacted indices bit-exact.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cases
from oracle import slcl_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(autouse=True, scope="module")
def _register_ops():
    """torch.ops.slcl.* exists once slcl.ops has been imported: tests that call the ops directly must not depend on an
    earlier test having imported the package (e.g. under ``-k``)."""
    import slcl.ops  # noqa: F401


def dev():
    return torch.device("cuda:0")


def close(a, b, rtol=RTOL, atol=1e-6):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def grad_close(a, b, rtol=RTOL, floor=0.1):
    """Gradients: rtol relative per element, with an absolute floor of floor*rtol*max|grad| for the
    elements near zero (sums with cancellation carry the absolute error of the large terms)."""
    b = torch.as_tensor(b).double()
    scale = float(b.abs().max()) + 1e-30
    close(a, b, rtol=rtol, atol=rtol * scale * floor)


@pytest.fixture(scope="module")
def api():
    from slcl import loss, utils_
    return loss, utils_


# ---------------------------------------------------------------------------
# prototype path
# ---------------------------------------------------------------------------
def test_kat1_source_loss_vs_reference_golden(api, golden):
    loss_mod, _ = api
    feas, labels = cases.kat1()
    cc = cases.shipped_centres()
    f = feas.to(dev()).requires_grad_(True)
    c = cc.to(dev()).requires_grad_(True)
    mp = loss_mod.MPCL(dev(), num_class=4, temperature=.1, base_temperature=1, m=.4)
    out = loss_mod.mpcl_loss_calc(f, labels.to(dev()), c, mp, tag='source')
    out.backward()
    assert abs(out.item() - 0.23589777946472168) < 0.23589777946472168 * RTOL
    close(out, golden["kat1_loss"])
    grad_close(f.grad, golden["kat1_dfeas"])
    grad_close(c.grad, golden["kat1_dcentres"])


def test_kat2_pseudo_label_and_target_loss(api, golden):
    loss_mod, utils_mod = api
    ft = cases.kat2().to(dev())
    cc = cases.shipped_centres().to(dev())
    hard, sel = utils_mod.generate_pseudo_label(ft, cc, .25)
    assert hard.dtype == torch.int64 and sel.dtype == torch.float32
    assert np.array_equal(hard.cpu().numpy(), golden["kat2_label"])            # bit-exact
    assert np.array_equal(sel.cpu().numpy(), golden["kat2_sel"])
    f = ft.clone().requires_grad_(True)
    mp = loss_mod.MPCL(dev(), num_class=4, temperature=.1, base_temperature=1, m=.2)
    out = loss_mod.mpcl_loss_calc(f, hard, cc, mp, pixel_sel_loc=sel, tag='target')
    out.backward()
    close(out, golden["kat2_loss"], atol=1e-8)
    grad_close(f.grad, golden["kat2_dfeas"])


def test_kat8_direct_mpcl_forward_soft_mask(api, golden):
    loss_mod, _ = api
    feas, _ = cases.kat1()
    cc = cases.shipped_centres()
    unit = F.normalize(feas, p=2, dim=1).permute(0, 2, 3, 1).reshape(-1, 32).to(dev()).requires_grad_(True)
    cen = F.normalize(cc, p=2, dim=1).t().to(dev())
    mp = loss_mod.MPCL(dev(), num_class=4, temperature=.1, base_temperature=1, m=.4)
    out = mp(unit.unsqueeze(1), None, cen, mask=cases.kat8_mask().to(dev()))
    out.backward()
    close(out, golden["kat8_loss"])
    grad_close(unit.grad, golden["kat8_dunit"])


@pytest.mark.parametrize("n,c,k,m,with_sel", [(300, 32, 4, .4, True), (257, 48, 5, .2, False), (64, 128, 3, .3, True)])
def test_mpcl_gradients_into_soft_mask_and_pixel_sel_loc(api, n, c, k, m, with_sel):
    """utils/loss.py:516-517 / :558-565 take `mask` and `pixel_sel_loc` as tensors; autograd of the reference's formula
    (the restatement) gives their gradients, slcl_proto_bwd_aux must match -- next to the usual d/d features."""
    loss_mod, _ = api
    gen = cases.g(300 + n)
    unit = F.normalize(torch.randn(n, c, generator=gen), dim=1)
    cen = F.normalize(torch.randn(k, c, generator=gen), dim=1).t().contiguous()
    mask = torch.softmax(2 * torch.randn(n, k, generator=gen), dim=1)
    sel = torch.rand(n, generator=gen) if with_sel else None
    spec = O.MarginSpec(num_class=k, temperature=.1, m=m, base_temperature=1.0)
    uo, mo = unit.clone().requires_grad_(True), mask.clone().requires_grad_(True)
    so = sel.clone().requires_grad_(True) if with_sel else None
    ref = O.mpcl_forward(spec, uo.unsqueeze(1), None, cen, pixel_sel_loc=so, mask=mo)
    ref.backward()
    ud, md = unit.to(dev()).requires_grad_(True), mask.to(dev()).requires_grad_(True)
    sd = sel.to(dev()).requires_grad_(True) if with_sel else None
    mp = loss_mod.MPCL(dev(), num_class=k, temperature=.1, base_temperature=1, m=m)
    out = mp(ud.unsqueeze(1), None, cen.to(dev()), pixel_sel_loc=sd, mask=md)
    out.backward()
    close(out, ref)
    grad_close(ud.grad, uo.grad)
    grad_close(md.grad, mo.grad)
    if with_sel:
        grad_close(sd.grad, so.grad)
    # hard labels: only pixel_sel_loc can ask for a gradient
    if with_sel:
        lab = torch.randint(0, k, (n,), generator=gen)
        so2 = sel.clone().requires_grad_(True)
        ref2 = O.mpcl_forward(spec, unit.unsqueeze(1), lab, cen, pixel_sel_loc=so2)
        ref2.backward()
        sd2 = sel.to(dev()).requires_grad_(True)
        out2 = mp(unit.to(dev()).unsqueeze(1), lab.to(dev()), cen.to(dev()), pixel_sel_loc=sd2)
        out2.backward()
        close(out2, ref2)
        grad_close(sd2.grad, so2.grad)


def test_kat9_cfg1_geometry_label_downsample(api, golden):
    """33x33 map (HW = 1089, odd: scalar path), labels 256 -> 33, K = 5, C = 128."""
    loss_mod, _ = api
    f9, lab9, cc9 = cases.kat9()
    f = f9.to(dev()).requires_grad_(True)
    mp = loss_mod.MPCL(dev(), num_class=5, temperature=.1, base_temperature=1, m=.4)
    out = loss_mod.mpcl_loss_calc(f, lab9.to(dev()), cc9.to(dev()), mp, tag='source')
    out.backward()
    close(out, golden["kat9_loss"])
    close(f.grad.abs().sum(), golden["kat9_dfeas_abs_sum"], rtol=RTOL)
    grad_close(f.grad[:, :4, :3, :3], golden["kat9_dfeas_head"])


def test_ragged_out_of_range_selection_easy_margin(api, golden):
    loss_mod, utils_mod = api
    rf, rl, rc, rsel = cases.ragged_case()
    f = rf.to(dev()).requires_grad_(True)
    c = rc.to(dev()).requires_grad_(True)
    mp = loss_mod.MPCL(dev(), num_class=5, temperature=.07, base_temperature=.07, m=.5)
    out = loss_mod.mpcl_loss_calc(f, rl.view(-1).to(dev()), c, mp, pixel_sel_loc=rsel.to(dev()), tag='target')
    out.backward()
    close(out, golden["ragged_loss"])
    grad_close(f.grad, golden["ragged_dfeas"])
    grad_close(c.grad, golden["ragged_dcentres"])
    hard, sel = utils_mod.generate_pseudo_label(rf.to(dev()), rc.to(dev()), .1)
    assert np.array_equal(hard.cpu().numpy(), golden["ragged_label"])
    assert np.array_equal(sel.cpu().numpy(), golden["ragged_sel"])
    f = rf.to(dev()).requires_grad_(True)
    mp = loss_mod.MPCL(dev(), num_class=5, temperature=.1, base_temperature=1, m=.4, easy_margin=True)
    out = loss_mod.mpcl_loss_calc(f, rl.to(dev()), rc.to(dev()), mp, tag='source')
    out.backward()
    close(out, golden["ragged_easy_loss"])
    grad_close(f.grad, golden["ragged_easy_dfeas"])


@pytest.mark.parametrize("b,c,h,w,k,with_sel", [
    (2, 32, 16, 16, 4, False),      # vector path, DRUNet channel count
    (3, 128, 8, 12, 5, True),       # vector path, cfg2 channel count, selection mask
    (2, 7, 5, 5, 3, True),          # scalar path, C not a multiple of the unroll
    (1, 40, 4, 4, 8, False),        # K = 8 (two broadcast loads per channel)
    (2, 16, 6, 6, 2, False),        # K = 2
])
def test_proto_loss_vs_oracle(api, b, c, h, w, k, with_sel):
    loss_mod, _ = api
    gen = cases.g(100 + c + k)
    feas = torch.randn(b, c, h, w, generator=gen) * 2.0
    labels = torch.randint(0, k, (b, h, w), generator=gen)
    cc = torch.randn(k, c, generator=gen)
    sel = (torch.rand(b * h * w, generator=gen) > 0.5).float() if with_sel else None
    spec = O.MarginSpec(num_class=k, temperature=.1, m=.4, base_temperature=1.0)
    fo = feas.clone().requires_grad_(True)
    co = cc.clone().requires_grad_(True)
    ref = O.mpcl_loss_calc(fo, labels if not with_sel else labels.view(-1), co, spec, pixel_sel_loc=sel,
                           tag='target' if with_sel else 'source')
    ref.backward()
    f = feas.to(dev()).requires_grad_(True)
    cg = cc.to(dev()).requires_grad_(True)
    mp = loss_mod.MPCL(dev(), num_class=k, temperature=.1, base_temperature=1, m=.4)
    out = loss_mod.mpcl_loss_calc(f, (labels if not with_sel else labels.view(-1)).to(dev()), cg, mp,
                                  pixel_sel_loc=None if sel is None else sel.to(dev()),
                                  tag='target' if with_sel else 'source')
    (out * 3.0).backward()            # non-unit upstream gradient
    close(out, ref)
    grad_close(f.grad / 3.0, fo.grad)
    grad_close(cg.grad / 3.0, co.grad)


def test_proto_loss_non_contiguous_batch_slice(api):
    """The trainers slice the batch (dcdr_ft[:s_size], Trainer_MCCL.py:271-273) and may hand over
    channel slices; both collapse to a strided map without a copy."""
    loss_mod, _ = api
    gen = cases.g(7)
    big = torch.randn(4, 24, 8, 8, generator=gen)
    labels = torch.randint(0, 4, (2, 8, 8), generator=gen)
    cc = torch.randn(4, 16, generator=gen)
    spec = O.MarginSpec(4, .1, .4, 1.0)
    view_cpu = big[1:3, 4:20]
    fo = view_cpu.clone().requires_grad_(True)
    ref = O.mpcl_loss_calc(fo, labels, cc, spec)
    ref.backward()
    bigg = big.to(dev()).requires_grad_(True)
    out = loss_mod.mpcl_loss_calc(bigg[1:3, 4:20], labels.to(dev()), cc.to(dev()),
                                  loss_mod.MPCL(dev(), 4, .1, .4, 1.0))
    out.backward()
    close(out, ref)
    grad_close(bigg.grad[1:3, 4:20], fo.grad)
    assert float(bigg.grad[0].abs().sum()) == 0.0 and float(bigg.grad[:, :4].abs().sum()) == 0.0


def test_kat3_ema_class_centres(api, golden):
    _, utils_mod = api
    feas, labels = cases.kat1()
    cc = cases.shipped_centres().to(dev())
    new = utils_mod.update_class_center_iter(feas.to(dev()), labels.to(dev()), cc, m=.9)
    close(new, golden["kat3_centres"])
    lab = labels.clone()
    lab[lab == 2] = 1
    new = utils_mod.update_class_center_iter(feas.to(dev()), lab.to(dev()), cc, m=.9)
    close(new, golden["kat3_empty_centres"])
    rf, rl, rc, _ = cases.ragged_case()
    new = utils_mod.update_class_center_iter(rf.to(dev()), rl.to(dev()), rc.to(dev()), m=.8, num_class=5)
    close(new, golden["ragged_ema"])


def test_class_counts_bit_exact():
    """Weight-sum column of the class sums = exact integer class counts."""
    gen = cases.g(55)
    feas = torch.randn(3, 32, 20, 20, generator=gen)
    labels = torch.randint(-1, 5, (3, 20, 20), generator=gen)       # includes out-of-range -1 and 4
    sums = torch.ops.slcl.class_sums(feas.to(dev()), labels.view(-1).to(dev()), None, False, 0.0, None, 1, 4)
    counts = sums[:, -1].cpu()
    expect = torch.stack([(labels == k).sum() for k in range(4)]).double()
    assert torch.equal(counts, expect)
    ref_sums = torch.stack([(feas * (labels == k).unsqueeze(1)).sum(dim=(0, 2, 3)) for k in range(4)])
    close(sums[:, :-1], ref_sums, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("b,c,h,w,k,parts,mode", [
    (3, 128, 24, 20, 5, 1, "hard"),        # two channel blocks (C = 128 > 64), pixel tail (480 = 3*128 + 96)
    (2, 32, 16, 16, 4, 2, "soft"),         # cfg5-like: CPW = 2, partitions
    (5, 12, 8, 12, 3, 1, "hard"),          # C < 16: CPW = 1, channel tail inside the TMA box
    (2, 72, 12, 12, 8, 2, "soft"),         # 16 weight columns (KWT = 16 -> CPW = 2), C tail in the last channel block
    (4, 64, 10, 10, 5, 3, "argmax"),       # arg-max one-hot with certainty threshold, 15 columns
    (1, 40, 33, 33, 5, 1, "hard"),         # HW % 4 != 0: the v2 kernel
])
def test_class_sums_kernels_vs_torch(b, c, h, w, k, parts, mode, monkeypatch):
    """slcl::class_sums against a plain fp64 torch evaluation, through the TMA-fed v3 kernel (default where the
    shape allows) and through the v2 kernel (SLCL_CLASS_SUMS_V2=1); weight sums of hard labels are exact integers."""
    gen = cases.g(700 + c + h)
    feat = torch.randn(b, c, h, w, generator=gen)
    n = b * h * w
    part = (torch.randperm(n, generator=gen) % parts).to(torch.int32) if parts > 1 else None
    pid = part.long() if part is not None else torch.zeros(n, dtype=torch.long)
    if mode == "hard":
        labels = torch.randint(-1, k + 1, (n,), generator=gen)           # includes out-of-range labels
        wts = torch.zeros(n, parts * k, dtype=torch.float64)
        ok = (labels >= 0) & (labels < k)
        wts[torch.arange(n)[ok], (pid * k + labels)[ok]] = 1.0
        args = (labels.to(dev()), None, False, 0.0)
    else:
        probs = torch.softmax(3 * torch.randn(b, k, h, w, generator=gen), dim=1)
        pf = probs.permute(0, 2, 3, 1).reshape(n, k).double()
        thr = 0.6 if mode == "argmax" else 0.0
        cert = (pf.max(1).values >= thr).double() if thr > 0 else torch.ones(n, dtype=torch.float64)
        rows = pf * cert[:, None] if mode == "soft" else torch.nn.functional.one_hot(pf.argmax(1), k).double() * cert[:, None]
        wts = torch.zeros(n, parts * k, dtype=torch.float64)
        for pp in range(parts):
            wts[pid == pp, pp * k:(pp + 1) * k] = rows[pid == pp]
        args = (None, probs.to(dev()), mode == "soft", thr)
    x = feat.permute(0, 2, 3, 1).reshape(n, c).double()
    ref = torch.cat([wts.t() @ x, wts.sum(0)[:, None]], dim=1)
    for v2 in ("0", "1"):
        monkeypatch.setenv("SLCL_CLASS_SUMS_V2", v2)
        got = torch.ops.slcl.class_sums(feat.to(dev()), args[0], args[1], args[2], args[3],
                                        part.to(dev()) if part is not None else None, parts, k).cpu()
        close(got[:, :-1], ref[:, :-1], rtol=1e-5, atol=1e-4)
        if mode == "hard":
            assert torch.equal(got[:, -1], ref[:, -1])
        else:
            close(got[:, -1], ref[:, -1], rtol=1e-5)


# ---------------------------------------------------------------------------
# centroid path
# ---------------------------------------------------------------------------
def test_kat5_hard_centroids(api, golden):
    _, utils_mod = api
    feas, labels = cases.kat1()
    ft = cases.kat2()
    c1, ratio, std = utils_mod.cal_centroid(feas.to(dev()), labels.to(dev()), momentum=.9)
    assert ratio is None and std == []
    close(c1, golden["kat5_c1"])
    f = ft.to(dev()).requires_grad_(True)
    c2, _, _ = utils_mod.cal_centroid(f, labels.to(dev()), previous_centroid=c1.detach(), momentum=.9)
    (c2 * c2).sum().backward()
    close(c2, golden["kat5_c2"])
    grad_close(f.grad, golden["kat5_dft"])


@pytest.mark.parametrize("name,kw", [
    ("wtd", dict(weighted_ave=True)),
    ("wtd_thd", dict(weighted_ave=True, threshold=0.6)),
    ("hardpl_thd", dict(weighted_ave=False, threshold=0.6)),
    ("wtd_ema", dict(weighted_ave=True, momentum=.9)),
])
def test_soft_centroids_and_contrastive_vs_repaired_reference(api, golden, name, kw):
    loss_mod, utils_mod = api
    sft, probs = cases.soft_case()
    if name == "wtd_ema":
        kw = dict(kw, previous_centroid=cases.shipped_centres().to(dev()))
    f = sft.to(dev()).requires_grad_(True)
    p = probs.to(dev()).requires_grad_(True)
    cen, _, _ = utils_mod.cal_centroid(f, p, pseudo_label=True, **kw)
    src_c, _ = cases.kat4()
    out = loss_mod.ContrastiveLoss()(src_c.to(dev()), cen) + (cen * cen).sum()
    out.backward()
    close(cen, golden[f"soft_{name}_cen"])
    close(out, golden[f"soft_{name}_loss"])
    grad_close(f.grad, golden[f"soft_{name}_dft"])
    if golden[f"soft_{name}_dp"].size:
        grad_close(p.grad, golden[f"soft_{name}_dp"])
    else:
        assert p.grad is None or float(p.grad.abs().sum()) == 0.0


@pytest.mark.parametrize("k,parts,thr", [(4, 2, None), (5, 2, 0.5), (4, 3, None), (8, 2, None)])
def test_rmc_partitions_vs_oracle(api, k, parts, thr):
    """Partitions are our spec (parity unpinned by the reference): part_id comes from PyTorch's RNG
    stream on the host, so both sides consume identical indices; per-partition weight sums match."""
    _, utils_mod = api
    ft, probs = cases.soft_case(seed=40 + k, b=2, c=24, h=10, w=12, k=k)
    pid = O.rmc_partition_ids(2 * 10 * 12, parts, cases.g(9))
    fo = ft.clone().requires_grad_(True)
    po = probs.clone().requires_grad_(True)
    ref, _, _ = O.cal_centroid(fo, po, pseudo_label=True, weighted_ave=True, n_class=k, partition=parts,
                               threshold=thr, part_id=pid)
    sum((r * (i + 1.5)).pow(2).sum() for i, r in enumerate(ref)).backward()
    f = ft.to(dev()).requires_grad_(True)
    p = probs.to(dev()).requires_grad_(True)
    got, _, _ = utils_mod.cal_centroid(f, p, pseudo_label=True, weighted_ave=True, n_class=k, partition=parts,
                                       threshold=thr, part_id=pid.to(dev()))
    assert isinstance(got, list) and len(got) == parts
    sum((r * (i + 1.5)).pow(2).sum() for i, r in enumerate(got)).backward()
    for a, b in zip(got, ref):
        close(a, b, atol=1e-6)
    grad_close(f.grad, fo.grad)
    grad_close(p.grad, po.grad)
    # same generator state -> same partition ids as the product's own sampler
    mine = utils_mod.rmc_partition_ids(240, parts, cases.g(9))
    assert torch.equal(mine, pid)


def test_kat4_contrastive_loss_and_cnr(api, golden):
    loss_mod, _ = api
    cs, ct = cases.kat4()
    for name, kw in (("plain", {}), ("split", {"split": True}), ("bg", {"bg": True})):
        a = cs.to(dev()).requires_grad_(True)
        b = ct.to(dev()).requires_grad_(True)
        out = loss_mod.ContrastiveLoss(tau=5)(a, b, **kw)
        (out * 2.0).backward()
        close(out, golden[f"kat4_{name}_loss"])
        grad_close(a.grad / 2.0, golden[f"kat4_{name}_ds"])
        grad_close(b.grad / 2.0, golden[f"kat4_{name}_dt"])
    # tau is ignored by the reference (bug-compatible)
    assert loss_mod.ContrastiveLoss(tau=.1)(cs.to(dev()), ct.to(dev())).item() == \
        loss_mod.ContrastiveLoss(tau=5)(cs.to(dev()), ct.to(dev())).item()
    # CNR vs oracle
    so = cs.clone().requires_grad_(True)
    to = [ct.clone().requires_grad_(True), (ct * 1.7).clone().requires_grad_(True)]
    ref = O.cnr_loss(so, to)
    ref.backward()
    sg = cs.to(dev()).requires_grad_(True)
    tg = [ct.to(dev()).requires_grad_(True), (ct * 1.7).to(dev()).requires_grad_(True)]
    out = loss_mod.cnr_loss(sg, tg)
    out.backward()
    close(out, ref)
    grad_close(sg.grad, so.grad)
    grad_close(tg[1].grad, to[1].grad)


# ---------------------------------------------------------------------------
# sampler
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("parts,split,with_aug,k,c", [(2, False, True, 4, 32), (1, True, True, 4, 32), (3, False, False, 4, 48),
                                                      (2, True, True, 5, 128)])
def test_fused_mccl_centroid_losses_equal_the_drop_in_calls(api, parts, split, with_aug, k, c):
    """mccl_centroid_losses (one op, two launches) == the trainer's own loop over ContrastiveLoss / CNR
    (trainer/Trainer_MCCL.py:303-326) built from the drop-in modules, values and gradients."""
    loss_mod, _ = api
    g = cases.g(parts * 10 + k)
    cs = torch.randn(k, c, generator=g).to(dev())
    cts = [torch.randn(k, c, generator=g).to(dev()) for _ in range(parts)]
    ca = torch.randn(k, c, generator=g).to(dev()) if with_aug else None
    crit = loss_mod.ContrastiveLoss()

    def leaves():
        return (cs.clone().requires_grad_(True), [t.clone().requires_grad_(True) for t in cts],
                ca.clone().requires_grad_(True) if with_aug else None)
    s1, t1, a1 = leaves()
    inter = sum(crit(s1, t, split=split) for t in t1) / parts
    intra = sum(crit(t, a1, split=split) for t in t1) / parts if with_aug else 0.0
    ref = 0.7 * inter + 0.3 * intra + 4e-5 * loss_mod.cnr_loss(s1, t1)
    ref.backward()
    s2, t2, a2 = leaves()
    out, terms = loss_mod.mccl_centroid_losses(s2, t2, a2, inter_w=0.7, intra_w=0.3, cnr_w=4e-5, split=split)
    out.backward()
    close(out, ref, rtol=1e-5)
    close(terms[0], inter, rtol=1e-5)
    grad_close(s2.grad, s1.grad, rtol=1e-5)
    for x, y in zip(t2, t1):
        grad_close(x.grad, y.grad, rtol=1e-5)
    if with_aug:
        grad_close(a2.grad, a1.grad, rtol=1e-5)


@pytest.mark.parametrize("n,k", [(1, 4), (255, 4), (4096, 5), (4097, 5), (100003, 8)])
def test_compaction_bit_exact_with_nonzero(n, k):
    gen = cases.g(n + k)
    labels = torch.randint(-1, k + 1, (n,), generator=gen)       # out-of-range labels are dropped
    counts, offsets, index = torch.ops.slcl.compact_by_class(labels.to(dev()), k)
    counts, offsets, index = counts.cpu(), offsets.cpu(), index.cpu()
    for c in range(k):
        want = torch.nonzero(labels == c).squeeze(1)
        assert counts[c].item() == want.numel()
        assert torch.equal(index[offsets[c]:offsets[c + 1]], want)
    assert offsets[0].item() == 0 and offsets[k].item() == int(((labels >= 0) & (labels < k)).sum())


def test_gather_unit_rows_vs_oracle():
    gen = cases.g(77)
    feat = torch.randn(2, 40, 9, 7, generator=gen)
    idx = torch.randperm(2 * 9 * 7, generator=gen)[:50]
    want = O.gather_unit_rows(feat, idx)
    bf, f32, inv = torch.ops.slcl.gather_unit_rows(feat.to(dev()), idx.to(dev()), True, True, True)
    close(f32, want, rtol=1e-5, atol=1e-7)
    assert bf.shape == (50, 64)                                           # C = 40 padded to the 64-element swizzle row
    assert torch.equal(bf.cpu()[:, :40], f32.cpu().to(torch.bfloat16))   # bf16 rows are the rounded fp32 rows
    assert float(bf[:, 40:].abs().sum()) == 0.0                          # pad columns are zero
    rows = feat.permute(0, 2, 3, 1).reshape(-1, 40)[idx]
    close(inv, 1.0 / rows.norm(dim=1), rtol=1e-5)
    # backward through gather + normalise
    fo = feat.clone().requires_grad_(True)
    g_rows = torch.randn(50, 40, generator=gen)
    (O.gather_unit_rows(fo, idx) * g_rows).sum().backward()
    dfeat = torch.zeros_like(feat, device=dev())
    torch.ops.slcl.scatter_rows_bwd(feat.to(dev()), idx.to(dev()), True, g_rows.to(dev()), inv, dfeat)
    grad_close(dfeat, fo.grad)


@pytest.mark.parametrize("normalize", [True, False])
def test_scatter_rows_by_map_equals_the_sparse_scatter(normalize):
    """slcl_scatter_rows_by_map (pixel-side backward of the gather: every element of dfeat written once, two row sets
    summed) against two slcl_scatter_rows_bwd launches into a zeroed map; the sets overlap like anchors / contrast rows."""
    from slcl import ops
    gen = cases.g(79)
    feat = torch.randn(3, 48, 10, 7, generator=gen).to(dev())
    n = 3 * 10 * 7
    perm = torch.randperm(n, generator=gen)
    idx_b, idx_a = perm[:120].to(dev()), perm[60:150].to(dev())                     # 60 pixels in both sets, 60 / 30 in one only
    _, _, inv_a = torch.ops.slcl.gather_unit_rows(feat, idx_a, True, True, False)
    _, _, inv_b = torch.ops.slcl.gather_unit_rows(feat, idx_b, True, True, False)
    d_a = torch.randn(idx_a.numel(), 48, generator=gen).to(dev())
    d_b = torch.randn(idx_b.numel(), 48, generator=gen).to(dev())
    want = torch.zeros_like(feat)
    torch.ops.slcl.scatter_rows_bwd(feat, idx_a, normalize, d_a, inv_a, want)
    torch.ops.slcl.scatter_rows_bwd(feat, idx_b, normalize, d_b, inv_b, want)
    selfcol, selfrow, tables = ops.self_maps_bounded(idx_a, idx_b, n)
    # the self maps agree with the sort-based construction, and the tables are the pixel -> row maps
    sc2, sr2 = ops.self_maps(idx_a, idx_b)
    assert torch.equal(selfcol, sc2) and torch.equal(selfrow, sr2)
    assert torch.equal(tables[0][idx_a].long(), torch.arange(idx_a.numel(), device=dev()))
    got = ops.scatter_rows_by_map(feat, normalize, tables[0], d_a, inv_a, tables[1], d_b, inv_b)
    grad_close(got, want)
    one = ops.scatter_rows_by_map(feat, normalize, tables[1], d_b, inv_b, None, None, None)
    want1 = torch.zeros_like(feat)
    torch.ops.slcl.scatter_rows_bwd(feat, idx_b, normalize, d_b, inv_b, want1)
    grad_close(one, want1)


def test_tile_weights_equal_the_torch_formula():
    """slcl_tile_weights: fg / (foreground of the tile) / (tiles with foreground), utils/loss.py:382-384 and :445-448."""
    from slcl import ops
    gen = cases.g(81)
    n_tiles, per = 7, 96
    lab = torch.randint(0, 4, (n_tiles * per,), generator=gen)
    lab[2 * per:3 * per] = 0                                              # one all-background tile
    meta = ops.pad_meta(lab.to(dev()), torch.arange(n_tiles * per, device=dev()))
    w, tile_fg = ops.tile_weights(meta, n_tiles * per, n_tiles, True)
    fg = (lab != 0).float().view(n_tiles, per)
    tf = fg.sum(1, keepdim=True)
    keep = (tf > 0).float()
    want = (fg / tf.clamp_min(1.0) * keep / keep.sum().clamp_min(1.0)).reshape(-1)
    torch.testing.assert_close(w.cpu(), want, rtol=1e-6, atol=0)
    assert torch.equal(tile_fg.cpu(), tf.reshape(-1))
    w1, _ = ops.tile_weights(meta, n_tiles * per, 1, False)               # one problem: fg / sum fg
    torch.testing.assert_close(w1.cpu(), (fg / fg.sum()).reshape(-1), rtol=1e-6, atol=0)
    zero = ops.pad_meta(torch.zeros(64, dtype=torch.int64, device=dev()), torch.arange(64, device=dev()))
    assert torch.isnan(ops.tile_weights(zero, 64, 1, False)[0]).all()     # SupConLoss: 0/0, as the reference
    assert float(ops.tile_weights(zero, 64, 1, True)[0].abs().sum()) == 0.0


def test_rows_meta_equals_pad_meta():
    from slcl import ops
    gen = cases.g(80)
    labels = torch.randint(-1, 6, (500,), generator=gen).to(dev())
    idx = torch.randperm(500, generator=gen)[:130].to(dev())
    assert torch.equal(ops.rows_meta(labels, idx), ops.pad_meta(labels[idx], idx))


def test_scatter_rows_wide_map_distinct_and_repeated_rows():
    """C = 256 (the warp-per-row kernel, rows held in registers): distinct pixels and pixels that occur several times."""
    gen = cases.g(78)
    feat = torch.randn(2, 256, 6, 5, generator=gen)
    for idx in (torch.randperm(60, generator=gen)[:37], torch.randint(0, 60, (90,), generator=gen)):
        fo = feat.clone().requires_grad_(True)
        g_rows = torch.randn(idx.numel(), 256, generator=gen)
        (O.gather_unit_rows(fo, idx) * g_rows).sum().backward()
        _, _, inv = torch.ops.slcl.gather_unit_rows(feat.to(dev()), idx.to(dev()), True, True, False)
        dfeat = torch.zeros_like(feat, device=dev())
        torch.ops.slcl.scatter_rows_bwd(feat.to(dev()), idx.to(dev()), True, g_rows.to(dev()), inv, dfeat)
        grad_close(dfeat, fo.grad)


# ---------------------------------------------------------------------------
# size-independent properties at full size (cfg2 / cfg4 shapes)
# ---------------------------------------------------------------------------
def test_full_size_properties_cfg4_shape(api):
    """cfg4 per-GPU shape B16 C32 224x224 K4: (i) counts are exact and sum to N, (ii) EMA with m=1 is the
    identity, (iii) the loss of a batch equals the pixel-weighted mean of the losses of its halves,
    (iv) gradient of the mean loss sums linearly, (v) pseudo-label of a centre-aligned map is exact."""
    loss_mod, utils_mod = api
    torch.manual_seed(0)
    b, c, h, w, k = 16, 32, 224, 224, 4
    feas = torch.randn(b, c, h, w, device=dev())
    labels = torch.randint(0, k, (b, h, w), device=dev())
    cc = torch.randn(k, c, device=dev())
    sums = torch.ops.slcl.class_sums(feas, labels.view(-1), None, False, 0.0, None, 1, k)
    assert int(sums[:, -1].sum().item()) == b * h * w
    assert torch.equal(sums[:, -1].long(), torch.bincount(labels.view(-1), minlength=k))
    assert torch.equal(utils_mod.update_class_center_iter(feas, labels, cc, m=1.0), cc)
    mp = loss_mod.MPCL(dev(), num_class=k, temperature=.1, base_temperature=1, m=.4)
    whole = loss_mod.mpcl_loss_calc(feas, labels, cc, mp)
    h1 = loss_mod.mpcl_loss_calc(feas[:8], labels[:8], cc, mp)
    h2 = loss_mod.mpcl_loss_calc(feas[8:], labels[8:], cc, mp)
    close(whole, 0.5 * (h1 + h2), rtol=1e-5)
    f = feas.clone().requires_grad_(True)
    loss_mod.mpcl_loss_calc(f, labels, cc, mp).backward()
    fh = feas[:8].clone().requires_grad_(True)
    loss_mod.mpcl_loss_calc(fh, labels[:8], cc, mp).backward()
    grad_close(f.grad[:8] * 2.0, fh.grad, rtol=1e-5)
    # rows of a normalised gradient are orthogonal to the pixel (d/dx of a function of x/|x|)
    radial = (f.grad * feas).sum(1).abs().max().item()
    assert radial < 1e-9
    # a map whose pixels are the centres themselves: label = class, gap = 1 - max off-diagonal cosine
    aligned = cc[labels].permute(0, 3, 1, 2).contiguous()
    hard, sel = utils_mod.generate_pseudo_label(aligned, cc, 0.0)
    assert torch.equal(hard, labels.view(-1)) and bool((sel == 1).all())


# ---------------------------------------------------------------------------
# FULL-SIZE parity against the oracle (BASELINE.json configs at their stated sizes).  The oracle is plain torch, so it
# runs on CUDA tensors too: evaluated here on the same GPU in fp32 with TF32 off -- the reference's own op sequence at the
# reference's own precision.  Class counts / labels bit-exact, everything else rtol 1e-4.
# ---------------------------------------------------------------------------
@pytest.fixture()
def no_tf32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _proto_full_size(loss_mod, feas, labels, cc, k, m, sel=None, tag='source'):
    spec = O.MarginSpec(num_class=k, temperature=.1, m=m, base_temperature=1.0)
    fo = feas.clone().requires_grad_(True)
    ref = O.mpcl_loss_calc(fo, labels, cc, spec, pixel_sel_loc=sel, tag=tag)
    ref.backward()
    f = feas.clone().requires_grad_(True)
    mp = loss_mod.MPCL(dev(), num_class=k, temperature=.1, base_temperature=1, m=m)
    out = loss_mod.mpcl_loss_calc(f, labels, cc, mp, pixel_sel_loc=sel, tag=tag)
    out.backward()
    close(out, ref, rtol=RTOL, atol=0)
    grad_close(f.grad, fo.grad)
    return out


def test_full_size_cfg1_batch8(api, no_tf32):
    """configs[0] at its stated size: B8, 128-d, 33x33 map, labels 256x256 (down-sampled inside), K5 -- CPU oracle."""
    loss_mod, _ = api
    g = cases.g(101)
    feas = torch.randn(8, 128, 33, 33, generator=g)
    labels = torch.randint(0, 5, (8, 256, 256), generator=g)
    cc = torch.randn(5, 128, generator=g)
    spec = O.MarginSpec(num_class=5, temperature=.1, m=.4, base_temperature=1.0)
    fo = feas.clone().requires_grad_(True)
    ref = O.mpcl_loss_calc(fo, labels, cc, spec, tag='source')
    ref.backward()
    f = feas.to(dev()).requires_grad_(True)
    out = loss_mod.mpcl_loss_calc(f, labels.to(dev()), cc.to(dev()), loss_mod.MPCL(dev(), 5, .1, .4, 1.0), tag='source')
    out.backward()
    close(out, ref, rtol=RTOL, atol=0)
    grad_close(f.grad, fo.grad)


def test_full_size_cfg2_prototype_path(api, no_tf32):
    """configs[1]: B32 C128 256x256 K5 (2 097 152 pixels), target variant with a selection mask: loss + dF vs the
    oracle on the GPU; EMA class centres and pseudo labels on the same map."""
    loss_mod, utils_mod = api
    gen = torch.Generator(device=dev()).manual_seed(202)
    b, c, h, w, k = 32, 128, 256, 256, 5
    feas = torch.randn(b, c, h, w, device=dev(), generator=gen)
    labels = torch.randint(0, k, (b * h * w,), device=dev(), generator=gen)
    sel = (torch.rand(b * h * w, device=dev(), generator=gen) > 0.5).float()
    cc = torch.randn(k, c, device=dev(), generator=gen)
    _proto_full_size(loss_mod, feas, labels, cc, k, .2, sel=sel, tag='target')
    new = utils_mod.update_class_center_iter(feas, labels.view(b, h, w), cc, m=.9, num_class=k)
    close(new, O.update_class_center_iter(feas, labels.view(b, h, w), cc, m=.9, num_class=k), rtol=RTOL)
    sums = torch.ops.slcl.class_sums(feas, labels, None, False, 0.0, None, 1, k)
    assert torch.equal(sums[:, -1].long(), torch.bincount(labels, minlength=k))


def _pseudo_labels_match(utils_mod, feas, cc, th):
    hard, sel = utils_mod.generate_pseudo_label(feas, cc, th)
    hard_o, sel_o = O.generate_pseudo_label(feas, cc, th)
    if torch.equal(hard, hard_o) and torch.equal(sel, sel_o):
        return
    # two fp32 evaluation orders can only disagree at a near-tie: every mismatch must sit within 1e-6 of the decision
    fn = F.normalize(feas, dim=1).permute(0, 2, 3, 1).reshape(-1, feas.shape[1]).double()
    cos = fn @ F.normalize(cc, dim=1).double().t()
    top = torch.sort(cos, dim=1).values
    gap = top[:, -1] - top[:, -2]
    bad_l = hard != hard_o
    bad_s = sel != sel_o
    assert int(bad_l.sum()) + int(bad_s.sum()) <= 4
    assert bool((gap[bad_l] < 1e-6).all()) and bool(((gap[bad_s] - th).abs() < 1e-6).all())


def test_full_size_cfg4_per_gpu_shape(api, no_tf32):
    """configs[3] per GPU: 16 x 32 x 224 x 224, K4 (MS-CMRSeg): source loss with labels at another resolution is not the
    trainer's case here -- labels at feature resolution, source + target variants, EMA centres, pseudo labels."""
    loss_mod, utils_mod = api
    gen = torch.Generator(device=dev()).manual_seed(404)
    b, c, h, w, k = 16, 32, 224, 224, 4
    feas = torch.randn(b, c, h, w, device=dev(), generator=gen)
    labels = torch.multinomial(torch.tensor([0.9146, 0.0253, 0.0309, 0.0292], device=dev()), b * h * w, True, generator=gen)
    cc = cases.shipped_centres().to(dev())
    _proto_full_size(loss_mod, feas, labels.view(b, h, w), cc, k, .4, tag='source')
    new = utils_mod.update_class_center_iter(feas, labels.view(b, h, w), cc, m=.9, num_class=k)
    close(new, O.update_class_center_iter(feas, labels.view(b, h, w), cc, m=.9, num_class=k), rtol=RTOL)
    _pseudo_labels_match(utils_mod, feas, new, .05)
    hard, sel = utils_mod.generate_pseudo_label(feas, new, .05)
    _proto_full_size(loss_mod, feas, hard, new, k, .2, sel=sel, tag='target')


def test_full_size_cfg5_soft_centroids_two_partitions(api, no_tf32):
    """configs[4] per GPU: 64 x 32 x 224 x 224 decoder features, K4 soft labels x certainty, P=2 reversed-Monte-Carlo
    partitions: centroids, dF and d(soft labels) vs the oracle; hard source centroids; per-partition class counts exact."""
    _, utils_mod = api
    gen = torch.Generator(device=dev()).manual_seed(505)
    b, c, h, w, k, parts = 64, 32, 224, 224, 4, 2
    n = b * h * w
    feas = torch.randn(b, c, h, w, device=dev(), generator=gen)
    probs = torch.softmax(3 * torch.randn(b, k, h, w, device=dev(), generator=gen), 1)
    part = (torch.randperm(n, device=dev(), generator=gen) % parts).to(torch.int32)
    lab = torch.randint(0, k, (b, h, w), device=dev(), generator=gen)
    gcen = torch.randn(parts * k, c, device=dev(), generator=gen)
    fo, po = feas.clone().requires_grad_(True), probs.clone().requires_grad_(True)
    ref, _, _ = O.cal_centroid(fo, po, pseudo_label=True, weighted_ave=True, partition=parts, n_class=k, threshold=.5,
                               part_id=part)
    (torch.cat(ref) * gcen).sum().backward()
    f, p = feas.clone().requires_grad_(True), probs.clone().requires_grad_(True)
    out, _, _ = utils_mod.cal_centroid(f, p, pseudo_label=True, weighted_ave=True, partition=parts, n_class=k, threshold=.5,
                                       part_id=part)
    (torch.cat(out) * gcen).sum().backward()
    close(torch.cat(out), torch.cat(ref), rtol=RTOL)
    grad_close(f.grad, fo.grad)
    grad_close(p.grad, po.grad)
    del fo, po, ref
    hard, _, _ = utils_mod.cal_centroid(feas, lab, n_class=k)
    hard_o, _, _ = O.cal_centroid(feas, lab, n_class=k)
    close(hard, hard_o, rtol=RTOL)
    # arg-max one-hot weights x partitions: the weight column of the class sums is an exact pixel count
    sums = torch.ops.slcl.class_sums(feas, None, probs, False, 0.0, part, parts, k)
    want = torch.bincount(part.long() * k + probs.argmax(1).reshape(-1), minlength=parts * k)
    assert torch.equal(sums[:, -1].long(), want)


# ---------------------------------------------------------------------------
# pixel <-> pixel path (tcgen05 tensor cores, bf16 inputs / fp32 accumulate): rtol 2e-2 (north_star)
# ---------------------------------------------------------------------------
P2P_RTOL = 2e-2


def test_kat6_supcon_vs_reference_golden(api, golden):
    loss_mod, _ = api
    f5, lab = cases.kat6()
    f = f5.to(dev()).requires_grad_(True)
    out = loss_mod.SupConLoss(.7)(f, lab.to(dev()))
    out.backward()
    close(out, golden["kat6_loss"], rtol=P2P_RTOL)
    assert abs(out.item() - 4.883881568908691) < 4.883881568908691 * 2e-3       # in practice ~1e-4
    grad_close(f.grad, golden["kat6_dfeat"], rtol=P2P_RTOL, floor=0.5)
    f = f5.to(dev()).requires_grad_(True)
    out = loss_mod.SupConLoss(.7)(f)                                              # unlabelled: other views are the positives
    out.backward()
    close(out, golden["kat6_unlab_loss"], rtol=P2P_RTOL)
    grad_close(f.grad, golden["kat6_unlab_dfeat"], rtol=P2P_RTOL, floor=0.5)
    from slcl import losses
    close(losses.SupConLoss(.7)(f5.to(dev()), lab.to(dev())), golden["kat6_dup_loss"], rtol=P2P_RTOL)
    # analytic sweeps (labels declared as class indices 0..3): same numbers
    f = f5.to(dev()).requires_grad_(True)
    out = loss_mod.SupConLoss(.7, n_class=4)(f, lab.to(dev()))
    out.backward()
    close(out, golden["kat6_loss"], rtol=P2P_RTOL)
    assert abs(out.item() - 4.883881568908691) < 4.883881568908691 * 2e-3
    grad_close(f.grad, golden["kat6_dfeat"], rtol=P2P_RTOL, floor=0.5)


def test_kat7_local_and_block_con_loss(api, golden):
    loss_mod, _ = api
    f7, lab7 = cases.kat7()
    f = f7.to(dev())
    close(loss_mod.LocalConLoss(.7, 4)(f, lab7.to(dev())), golden["kat7_local"], rtol=P2P_RTOL)
    close(loss_mod.BlockConLoss(.7, 32)(f, lab7.to(dev())), golden["kat7_block"], rtol=P2P_RTOL)
    close(loss_mod.LocalConLoss(.7, 4)(f), golden["kat7_local_unlab"], rtol=P2P_RTOL)
    close(loss_mod.LocalConLoss(.7, 4, n_class=4)(f, lab7.to(dev())), golden["kat7_local"], rtol=P2P_RTOL)
    close(loss_mod.BlockConLoss(.7, 32, n_class=4)(f, lab7.to(dev())), golden["kat7_block"], rtol=P2P_RTOL)
    # all-background labels: the reference returns 0 (early-out, utils/loss.py:405-407 / :445-447)
    zero = torch.zeros_like(lab7).to(dev())
    assert loss_mod.LocalConLoss(.7, 4)(f, zero).item() == 0.0
    assert loss_mod.BlockConLoss(.7, 32)(f, zero).item() == 0.0


@pytest.mark.parametrize("b,v,c,hw,bs,labelled", [
    (1, 2, 32, 64, 32, True),        # 4 tiles of 2048 rows: the block-diagonal batched sweeps
    (1, 2, 32, 224, 224 // 7, True), # 49 tiles of 2048 rows: BASELINE cfg3's map, more tiles than SMs / row tiles
    (1, 2, 16, 288, 8, True),        # 1296 tiles of 128 rows: more batches than finishing blocks (one table block per tile)
    (1, 2, 16, 64, 64, True),        # ONE tile of 8192 rows: n_batch = 1 with prebuilt metadata (no label sort in the general mode)
    (2, 2, 24, 48, 16, True),        # 9 tiles of 1024 rows, some all-background tiles
    (1, 2, 16, 64, 32, False),       # unlabelled: other views are the positives
    (1, 2, 16, 36, 12, True),        # 288 rows per tile (not a multiple of 128): the per-tile loop
])
def test_block_con_loss_vs_oracle(api, b, v, c, hw, bs, labelled):
    """BlockConLoss (utils/loss.py:416-466): all tiles as one block-diagonal problem (or the per-tile loop when a tile
    is not a whole number of 128-row tiles) against the oracle's per-tile evaluation, loss and gradient."""
    loss_mod, _ = api
    gen = cases.g(800 + hw + bs)
    f5 = F.normalize(torch.randn(b, v, c, hw, hw, generator=gen), dim=2)
    lab = torch.randint(0, 4, (b, v, hw, hw), generator=gen) if labelled else None
    if labelled and hw // bs > 1:
        lab[..., :bs, :bs] = 0                                          # one all-background tile: skipped (:439-440)
    fo = f5.clone().requires_grad_(True)
    ref = O.block_con_loss(fo, lab, 0.7, bs)
    ref.backward()
    # stock signature (class-index labels -> the analytic sweeps, one table of class sums per tile) and n_class=0
    # (any integer labels -> the general sweeps)
    for n_class in ((None, 0) if labelled else (None,)):
        f = f5.to(dev()).requires_grad_(True)
        out = loss_mod.BlockConLoss(0.7, bs, n_class=n_class)(f, lab.to(dev()) if labelled else None)
        out.backward()
        close(out, ref, rtol=P2P_RTOL)
        grad_close(f.grad, fo.grad, rtol=P2P_RTOL, floor=0.5)


def test_p2p_shift_is_the_row_norm_bound():
    """slcl_p2p_shift: shift_i = |a_i| max_j |b_j| / T from the gather's inv_norm outputs (one launch)."""
    from slcl import ops
    gen = cases.g(4242)
    for na, m in ((1, 1), (100, 37), (5000, 70001)):
        inv_a = (torch.rand(na, generator=gen) + 0.05).to(dev())
        inv_b = (torch.rand(m, generator=gen) + 0.05).to(dev())
        out = ops.p2p_shift(inv_a, inv_b, 0.7)
        ref = (1.0 / inv_a) * ((1.0 / inv_b).amax() / 0.7)
        torch.testing.assert_close(out, ref, rtol=1e-6, atol=0)


def test_supcon_unnormalised_features_vs_oracle(api):
    """SupConLoss does not normalise (utils/loss.py:342-349): rows of norm ~2, T = 0.5 -> the exp shift matters."""
    loss_mod, _ = api
    gen = cases.g(61)
    f5 = torch.randn(2, 2, 48, 6, 6, generator=gen) * 0.3
    lab = torch.randint(0, 3, (2, 2, 6, 6), generator=gen)
    fo = f5.clone().requires_grad_(True)
    ref = O.supcon_loss(fo, lab, 0.5)
    ref.backward()
    f = f5.to(dev()).requires_grad_(True)
    out = loss_mod.SupConLoss(0.5)(f, lab.to(dev()))
    out.backward()
    close(out, ref, rtol=P2P_RTOL)
    grad_close(f.grad, fo.grad, rtol=P2P_RTOL, floor=0.5)


@pytest.mark.parametrize("analytic", [True, False])
@pytest.mark.parametrize("b,c,h,w,k,na,nc,t", [
    (2, 40, 24, 24, 4, 200, 600, 0.7),        # ragged: A, M not multiples of the tiles; C padded 40 -> 64
    (2, 256, 16, 16, 5, 128, 500, 0.07),      # cfg3 dimensionality and temperature stress (500 of the 512 pixels)
    (1, 96, 12, 12, 3, 60, 144, 0.2),         # anchors == a third of all pixels
    (4, 128, 32, 32, 5, 1000, 3000, 0.7),     # several row tiles and column splits
])
def test_sampled_rectangular_loss_vs_oracle(b, c, h, w, k, na, nc, t, analytic):
    """Sampler + rectangular loss are our spec (parity unpinned by the reference); the indices come from
    PyTorch's RNG stream on both sides, so they must be bit-identical."""
    from slcl import p2p
    gen = cases.g(90 + c)
    feat = torch.randn(b, c, h, w, generator=gen)
    labels = torch.randint(0, k, (b, h, w), generator=gen)
    perm_o = torch.randperm(labels.numel(), generator=cases.g(5))          # the one draw, from PyTorch's RNG stream
    c_idx_o = O.sample_class_balanced(labels.view(-1), nc, k, perm_o)
    a_idx_o = O.sample_class_balanced(labels.view(-1), na, k, perm_o)
    a_idx, c_idx = p2p.sample_class_balanced(labels.to(dev()), na, nc, k, cases.g(5))
    assert torch.equal(a_idx.cpu(), a_idx_o) and torch.equal(c_idx.cpu(), c_idx_o)        # bit-exact selection
    fo = feat.clone().requires_grad_(True)
    lab_flat = labels.view(-1)
    ref = O.supcon_rect(O.gather_unit_rows(fo, a_idx_o), O.gather_unit_rows(fo, c_idx_o), lab_flat[a_idx_o],
                        lab_flat[c_idx_o], a_idx_o, c_idx_o, t)
    ref.backward()
    f = feat.to(dev()).requires_grad_(True)
    out = p2p.sampled_supcon_loss(f, labels.to(dev()), na, nc, k, temperature=t, anchor_idx=a_idx, contrast_idx=c_idx,
                                  analytic=analytic)
    out.backward()
    close(out, ref, rtol=P2P_RTOL)
    grad_close(f.grad, fo.grad, rtol=P2P_RTOL, floor=0.5)


@pytest.mark.parametrize("n,k,na,nc,rare", [(5000, 4, 64, 256, 0), (5000, 5, 100, 400, 7), (3000, 3, 30, 90, 1), (70000, 5, 4096, 16384, 0)])
def test_sampler_without_host_sync_bit_exact(n, k, na, nc, rare):
    """Sampler spec (oracle/slcl_oracle.py:sample_class_balanced): one randperm, rank within class from the stable
    compaction kernel, short classes topped up in permutation order, labels outside [0, K) never picked.  Indices must be
    bit-identical to the oracle; no .cpu()/.item() happens inside (checked with torch's sync debug mode)."""
    from slcl import p2p
    g = cases.g(300 + n + rare)
    labels = torch.randint(0, k, (n,), generator=g)
    if rare:
        labels[labels == k - 1] = 0
        labels[torch.randperm(n, generator=g)[:rare]] = k - 1       # class k-1 has only `rare` members: phase 2 fills its slots
    labels[torch.randperm(n, generator=g)[:17]] = 255                # ignore label: never picked
    dgen = lambda: torch.Generator(device=dev()).manual_seed(5)      # the draw itself runs on the device (torch's CUDA RNG)
    perm_o = torch.randperm(n, generator=dgen(), device=dev()).cpu()
    want_c = O.sample_class_balanced(labels, nc, k, perm_o)
    want_a = O.sample_class_balanced(labels, na, k, perm_o)
    lab_d = labels.to(dev())
    p2p.sample_class_balanced(lab_d, na, nc, k, dgen())               # warm-up (allocator)
    gen = dgen()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        a_idx, c_idx, a_fill, c_fill = p2p.sample_class_balanced(lab_d, na, nc, k, gen, return_counts=True)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert torch.equal(c_idx.cpu(), want_c) and torch.equal(a_idx.cpu(), want_a)
    assert int(c_fill) == want_c.numel() and int(a_fill) == want_a.numel()
    assert c_idx.unique().numel() == c_idx.numel()                    # distinct pixels
    assert bool((labels[c_idx.cpu()] < k).all())


def test_sampled_supcon_loss_whole_call_has_no_host_sync():
    """draw + compaction + self maps + gather + sweeps + backward scatter: nothing in the public cfg3 call may synchronise
    (torch's sync debug mode raises on .item() / .cpu() / boolean-mask indexing / nonzero)."""
    from slcl import p2p
    g = cases.g(4343)
    feat = torch.randn(2, 48, 32, 32, generator=g).to(dev())
    labels = torch.randint(0, 5, (2, 32, 32), generator=g).to(dev())
    gen = torch.Generator(device=dev()).manual_seed(11)

    def step():
        f = feat.clone().requires_grad_(True)
        out = p2p.sampled_supcon_loss(f, labels, 256, 1024, 5, temperature=0.7, generator=gen)
        out.backward()
        return out.detach(), f.grad
    step()                                            # warm-up: allocator, lazy module state
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        out, grad = step()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert torch.isfinite(out) and torch.isfinite(grad).all()


def test_sampler_reports_short_maps_and_the_loss_is_nan():
    from slcl import p2p
    g = cases.g(77)
    feat = torch.randn(1, 16, 8, 8, generator=g).to(dev())
    labels = torch.randint(0, 4, (1, 8, 8), generator=g).to(dev())
    a_idx, c_idx, a_fill, c_fill = p2p.sample_class_balanced(labels, 32, 128, 4, cases.g(1), return_counts=True)
    assert int(c_fill) == 64 and c_idx.numel() == 128                  # only 64 pixels exist
    assert torch.isnan(p2p.sampled_supcon_loss(feat, labels, 32, 128, 4, generator=cases.g(1)))


def test_p2p_full_size_properties_cfg3():
    """cfg3 size (4096 anchors x 16384 contrast rows, d = 256): loss is invariant to the exp shift, equals a
    chunked fp32 torch evaluation on the same bf16 rows, and duplicating every contrast row (with fresh ids)
    adds exactly log-sum-exp consistent terms (Z doubles for non-self rows)."""
    op = torch.ops.slcl
    g = torch.Generator(device=dev()).manual_seed(3)
    na, m, d, t = 4096, 16384, 256, 0.7
    b = F.normalize(torch.randn(m, d, device=dev(), generator=g), dim=1).to(torch.bfloat16)
    lb = torch.randint(0, 5, (m,), device=dev(), generator=g, dtype=torch.int32)
    ib = torch.arange(m, device=dev(), dtype=torch.int32)
    pick = torch.randperm(m, device=dev(), generator=g)[:na]
    a, la, ia = b[pick].contiguous(), lb[pick].contiguous(), ib[pick].contiguous()
    fg = (la != 0).float()
    w = fg / fg.sum()
    from slcl import ops as slcl_ops
    ma, mb = slcl_ops.pad_meta(la, ia), slcl_ops.pad_meta(lb, ib)
    s1 = torch.full((na,), 1.0 / t, device=dev())
    l1, st1, _ = op.p2p_fwd(a, b, ma, mb, s1, w, t)
    l2, _, _ = op.p2p_fwd(a, b, ma, mb, s1 * 1.5 + 0.3, w, t)
    close(l1, l2, rtol=1e-5)
    s = a.float() @ b.float().t() / t
    notself = ia.view(-1, 1) != ib.view(1, -1)
    pos = (la.view(-1, 1) == lb.view(1, -1)) & notself
    lz = torch.logsumexp(s.masked_fill(~notself, float("-inf")), dim=1)
    row = lz - (s * pos).sum(1) / pos.sum(1)
    close(l1, (row * w).sum(), rtol=1e-4)
    assert torch.equal(st1[:, 2].long(), pos.sum(1))                       # positive counts are exact
    # analytic sweeps (class-index labels): same statistics, loss and gradients as the per-element path and as
    # fp32 torch autograd on the same bf16 rows
    selfcol, selfrow = slcl_ops.self_maps(ia, ib)
    assert torch.equal(selfcol.long(), pick) and torch.equal(selfrow[pick].long(), torch.arange(na, device=dev()))
    l3, st3, state = op.p2p_fwd(a, b, ma, mb, s1, w, t, 5, selfcol, True)
    l4, st4, _ = op.p2p_fwd(a, b, ma, mb, s1, w, t, 5, selfcol, False)
    close(l3, l1, rtol=1e-5)
    close(l4, l1, rtol=1e-5)
    assert torch.equal(st3[:, 2].long(), pos.sum(1))
    assert torch.allclose(st3[:, 0], st1[:, 0], rtol=1e-4) and torch.allclose(st3[:, 1], st1[:, 1], rtol=1e-3, atol=1e-2)
    assert torch.equal(st3, st4)
    g_out = torch.full((1,), 0.7, device=dev())
    da_g, db_g = op.p2p_bwd(a, b, d, ma, mb, s1, w, t, st1, g_out, True, True)
    da_a, db_a = op.p2p_bwd(a, b, d, ma, mb, s1, w, t, st3, g_out, True, True, 5, selfcol, selfrow, state)
    da_r, db_r = op.p2p_bwd(a, b, d, ma, mb, s1, w, t, st3, g_out, True, True, 5, selfcol, selfrow)   # state regenerated
    af, bf = a.float().requires_grad_(True), b.float().requires_grad_(True)
    sf = af @ bf.t() / t
    lzf = torch.logsumexp(sf.masked_fill(~notself, float("-inf")), dim=1)
    ((lzf - (sf * pos).sum(1) / pos.sum(1)) * w).sum().mul(0.7).backward()
    for got_a, got_b in ((da_g, db_g), (da_a, db_a), (da_r, db_r)):
        grad_close(got_a, af.grad, rtol=P2P_RTOL, floor=0.5)
        grad_close(got_b, bf.grad, rtol=P2P_RTOL, floor=0.5)
    assert torch.equal(da_a, da_r) and torch.equal(db_a, db_r)             # deterministic, with or without the kept state


@pytest.mark.parametrize("na,m,d,k,t", [
    (9000, 9000, 192, 8, 0.5),      # > 8192 anchors: several anchors per finish warp; d = 192 (three 64-column chunks); K = 8
    (300, 5000, 64, 3, 0.2),        # one row tile, many column splits; d = 64 (ring smaller than the drain staging)
    (4100, 130, 128, 5, 1.0),       # row tail (4100 = 32*128 + 4), three column tiles (fewer than ring stages)
    (66000, 200, 64, 4, 0.7),       # > 65536 anchors: the beta~/ABsum pass walks several row blocks per CTA; most anchors
                                    # have no self pair
])
def test_p2p_analytic_equals_general_op_level(na, m, d, k, t):
    """Analytic sweeps (class sums outside the tensor-core sweep, forward keeps state) against the per-element general
    sweeps on the same bf16 rows: statistics, loss, dA, dB; anchors are contrast rows (self pairs) where na <= m."""
    from slcl import ops as slcl_ops
    op = torch.ops.slcl
    g = torch.Generator(device=dev()).manual_seed(100 + d)
    b = (F.normalize(torch.randn(m, d, device=dev(), generator=g), dim=1) * 1.3).to(torch.bfloat16)
    lb = torch.randint(0, k, (m,), device=dev(), generator=g, dtype=torch.int32)
    ib = torch.arange(m, device=dev(), dtype=torch.int32)
    if na <= m:
        pick = torch.randperm(m, device=dev(), generator=g)[:na]
        a, la, ia = b[pick].contiguous(), lb[pick].contiguous(), ib[pick].contiguous()
    else:                                                   # more anchors than contrast rows: the extra ones have no self pair
        extra = na - m
        a2 = (F.normalize(torch.randn(extra, d, device=dev(), generator=g), dim=1) * 0.9).to(torch.bfloat16)
        a = torch.cat([b, a2]).contiguous()
        la = torch.cat([lb, torch.randint(0, k, (extra,), device=dev(), generator=g, dtype=torch.int32)])
        ia = torch.cat([ib, torch.arange(m, m + extra, device=dev(), dtype=torch.int32)])
    w = torch.rand(na, device=dev(), generator=g)
    w = w / w.sum()
    ma, mb = slcl_ops.pad_meta(la, ia), slcl_ops.pad_meta(lb, ib)
    shift = (a.float().norm(dim=1) * b.float().norm(dim=1).max() / t).contiguous()
    selfcol, selfrow = slcl_ops.self_maps(ia, ib)
    l_g, st_g, _ = op.p2p_fwd(a, b, ma, mb, shift, w, t)
    l_a, st_a, state = op.p2p_fwd(a, b, ma, mb, shift, w, t, k, selfcol, True)
    close(l_a, l_g, rtol=1e-5)
    assert torch.equal(st_a[:, 2], st_g[:, 2])
    assert torch.allclose(st_a[:, 0], st_g[:, 0], rtol=2e-4) and torch.allclose(st_a[:, 1], st_g[:, 1], rtol=1e-3, atol=2e-2)
    g_out = torch.full((1,), -1.5, device=dev())            # negative upstream gradient
    da_g, db_g = op.p2p_bwd(a, b, d - 3, ma, mb, shift, w, t, st_g, g_out, True, True)
    da_a, db_a = op.p2p_bwd(a, b, d - 3, ma, mb, shift, w, t, st_a, g_out, True, True, k, selfcol, selfrow, state)
    grad_close(da_a, da_g, rtol=P2P_RTOL, floor=0.5)
    grad_close(db_a, db_g, rtol=P2P_RTOL, floor=0.5)
    only_b = op.p2p_bwd(a, b, d - 3, ma, mb, shift, w, t, st_a, g_out, False, True, k, selfcol, selfrow, state)[1]
    assert torch.equal(only_b, db_a)
    # general sweeps with self maps (no id tests, label-uniform tiles on the fast path): same again
    l_s, st_s, _ = op.p2p_fwd(a, b, ma, mb, shift, w, t, 0, selfcol)
    close(l_s, l_g, rtol=1e-5)
    assert torch.equal(st_s[:, 2], st_g[:, 2])
    assert torch.allclose(st_s[:, 0], st_g[:, 0], rtol=2e-4) and torch.allclose(st_s[:, 1], st_g[:, 1], rtol=1e-3, atol=2e-2)
    da_s, db_s = op.p2p_bwd(a, b, d - 3, ma, mb, shift, w, t, st_s, g_out, True, True, 0, selfcol, selfrow)
    grad_close(da_s, da_g, rtol=P2P_RTOL, floor=0.5)
    grad_close(db_s, db_g, rtol=P2P_RTOL, floor=0.5)
    # ... and with the contrast rows sorted by label (what slcl.p2p does for one-row-set problems): uniform tiles
    order = torch.argsort(lb.long(), stable=True)
    b2, lb2 = b[order].contiguous(), lb[order].contiguous()
    inv = torch.empty_like(order); inv[order] = torch.arange(m, device=dev())
    sc2 = torch.where(selfcol >= 0, inv[selfcol.clamp_min(0).long()].to(torch.int32), selfcol)
    sr2 = selfrow[order].contiguous()
    mb2 = slcl_ops.pad_meta(lb2, ib)
    l_o, st_o, _ = op.p2p_fwd(a, b2, ma, mb2, shift, w, t, 0, sc2.contiguous())
    close(l_o, l_g, rtol=1e-5)
    assert torch.equal(st_o[:, 2], st_g[:, 2])
    da_o, db_o = op.p2p_bwd(a, b2, d - 3, ma, mb2, shift, w, t, st_o, g_out, True, True, 0, sc2.contiguous(), sr2)
    grad_close(da_o, da_g, rtol=P2P_RTOL, floor=0.5)
    grad_close(db_o[inv], db_g, rtol=P2P_RTOL, floor=0.5)


def test_p2p_plan_graph_replay_matches_eager_launches():
    """slcl.plan.P2PPlan (raw C-ABI launches on fixed buffers): forward + backward captured into ONE CUDA graph (side-stream
    fork/join, programmatic dependent launches, the forward finish's ticket counter) and replayed several times gives bit
    for bit what the same launches give outside a graph; the loss also matches an fp32 torch evaluation."""
    from slcl import ops as slcl_ops
    from slcl.plan import P2PPlan
    g = torch.Generator(device=dev()).manual_seed(77)
    na, m, d, k, t = 1000, 3000, 128, 5, 0.7
    b = F.normalize(torch.randn(m, d, device=dev(), generator=g), dim=1).to(torch.bfloat16)
    lb = torch.randint(0, k, (m,), device=dev(), generator=g, dtype=torch.int32)
    ib = torch.arange(m, device=dev(), dtype=torch.int32)
    pick = torch.randperm(m, device=dev(), generator=g)[:na]
    a, la, ia = b[pick].contiguous(), lb[pick].contiguous(), ib[pick].contiguous()
    w = torch.rand(na, device=dev(), generator=g)
    w = (w / w.sum()).contiguous()
    shift = torch.full((na,), 1.0 / t, device=dev())
    selfcol, selfrow = slcl_ops.self_maps(ia, ib)
    plan = P2PPlan(a, b, d, slcl_ops.pad_meta(la, ia), slcl_ops.pad_meta(lb, ib), shift, w, t, n_class=k, a_selfcol=selfcol,
                   b_selfrow=selfrow)
    plan.forward(); plan.backward()
    torch.cuda.synchronize()
    loss0, da0, db0 = plan.loss.clone(), plan.d_a.clone(), plan.d_b.clone()
    af, bf = a.float(), b.float()
    sf = af @ bf.t() / t
    notself = ia.view(-1, 1) != ib.view(1, -1)
    pos = ((la.view(-1, 1) == lb.view(1, -1)) & notself).float()
    ref = ((torch.logsumexp(sf.masked_fill(~notself, float("-inf")), dim=1) - (sf * pos).sum(1) / pos.sum(1)) * w).sum()
    close(loss0[0], ref, rtol=P2P_RTOL)
    graph = plan.capture_graph()
    for _ in range(3):
        plan.loss.zero_(); plan.d_a.zero_(); plan.d_b.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(plan.loss, loss0) and torch.equal(plan.d_a, da0) and torch.equal(plan.d_b, db0)
    plan.forward(); plan.backward()                       # and eager launches still work after the graph used the counter
    torch.cuda.synchronize()
    assert torch.equal(plan.loss, loss0) and torch.equal(plan.d_a, da0)


def test_peer_exchange_kernel_single_rank_matches_rescale():
    """slcl_proto_rescale_peer with a world of one (the rank's own mailbox in ordinary device memory): the kernel stores its
    {epoch | fp32} words, finds them again, and must rewrite scal exactly as slcl_proto_rescale does -- over several calls,
    so both epoch parities and the slot reuse are exercised."""
    import ctypes as C
    from slcl import _lib
    from slcl._lib import check, ptr
    from slcl.peer import LoopbackMailboxes
    lib = _lib.load()
    assert lib.slcl_peer_mailbox_bytes(1, 2) // 8 == 8 + 4
    box = LoopbackMailboxes(1, dev(), capacity_words=2).boxes[0]
    peer = box.struct()
    stream = torch.cuda.current_stream().cuda_stream
    g = cases.g(5)
    for call in range(5):
        for has_sel in (0, 1):
            vals = torch.rand(4, generator=g) * 1000 + 1
            a, b = vals.to(dev()), vals.to(dev())
            check(lib.slcl_proto_rescale_peer(ptr(a), has_sel, C.byref(peer), stream), "slcl_proto_rescale_peer")
            check(lib.slcl_proto_rescale(ptr(b), has_sel, stream), "slcl_proto_rescale")
            torch.cuda.synchronize()
            assert torch.equal(a, b), (call, has_sel, a, b)
    assert box.epoch() == 10 and box.timeouts() == 0          # ten calls counted, no time-outs


# ---------------------------------------------------------------------------
# multi-rank protocol on ONE GPU: `world` ranks = `world` CUDA streams exchanging through loopback mailboxes
# (slcl.peer.LoopbackMailboxes).  The kernels, the mailbox protocol and the Python `group=` route are exactly what N
# processes over NVLink run; only the transport (same-device stores instead of P2P stores) differs.  bench.py repeats
# the sharded-vs-global checks on real multi-GPU boxes before it times anything.
# ---------------------------------------------------------------------------
def _rank_streams(world):
    return [torch.cuda.Stream(dev()) for _ in range(world)]


def _on_streams(streams, fn):
    """fn(rank) enqueued on streams[rank] for every rank, nothing synchronises in between; returns the results."""
    cur = torch.cuda.current_stream()
    out = []
    for r, st in enumerate(streams):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            out.append(fn(r))
    for st in streams:
        cur.wait_stream(st)
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_allreduce_f64_ranks_on_streams_bit_exact(world):
    from slcl.peer import LoopbackMailboxes
    op = torch.ops.slcl
    boxes = LoopbackMailboxes(world, dev(), capacity_words=2 * 700).boxes
    streams = _rank_streams(world)
    g = cases.g(77)
    for n in (1, 9, 700, 264):                       # several calls: both parities, slots reused with other sizes
        vals = [(torch.randn(n, generator=g, dtype=torch.float64) * 10 ** torch.randint(-3, 6, (n,), generator=g)).to(dev())
                for _ in range(world)]
        want = torch.zeros(n, dtype=torch.float64, device=dev())
        for v in vals:                               # rank order, like the kernel
            want = want + v
        bufs = [v.clone() for v in vals]
        torch.cuda.synchronize()
        _on_streams(streams, lambda r: op.peer_allreduce_f64(bufs[r], *boxes[r].args()))
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(bufs[r], want), (n, r)
    assert all(b.epoch() == 4 and b.timeouts() == 0 for b in boxes)


def test_peer_exchange_times_out_with_nan_instead_of_hanging():
    """A peer that never arrives: the waiting rank gives up after timeout_s, returns NaN and counts the event (ADVICE r1:
    the limit is a parameter now, default 10 minutes; PeerMailbox.check() turns the counter into an exception)."""
    from slcl.peer import LoopbackMailboxes
    box = LoopbackMailboxes(2, dev(), capacity_words=16, timeout_s=0.2).boxes[0]
    buf = torch.ones(3, dtype=torch.float64, device=dev())
    torch.ops.slcl.peer_allreduce_f64(buf, *box.args())          # rank 1 never calls
    torch.cuda.synchronize()
    assert torch.isnan(buf).all()
    assert box.timeouts() >= 1
    with pytest.raises(RuntimeError):
        box.check()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_class_centres_and_centroids_equal_global(api, world):
    """N sharded ranks must reproduce the 1-rank global result -- class counts bit-exactly -- for
    update_class_center_iter / cal_centroid(hard) / cal_centroid(soft, P=2) with the exchange inside the reduce kernel."""
    from slcl.peer import LoopbackMailboxes
    _, U = api
    op = torch.ops.slcl
    b, c, h, w, k, parts = 2 * world, 32, 24, 20, 4, 2
    g = cases.g(31)
    feat = torch.randn(b, c, h, w, generator=g).to(dev())
    lab = torch.randint(0, k, (b, h, w), generator=g)
    lab[lab == 3] = 1                                   # class 3 empty everywhere: the empty-class rule must see GLOBAL counts
    lab = lab.to(dev())
    probs = torch.softmax(3 * torch.randn(b, k, h, w, generator=g), 1).to(dev())
    part = (torch.randperm(b * h * w, generator=g) % parts).to(torch.int32).to(dev())
    cen = torch.randn(k, c, generator=g).to(dev())
    gcen = torch.randn(parts * k, c, generator=g).to(dev())
    boxes = LoopbackMailboxes(world, dev()).boxes
    streams = _rank_streams(world)
    per = b // world
    sl = lambda t, r: t[r * per:(r + 1) * per]
    pix = lambda t, r: t.reshape(b, -1)[r * per:(r + 1) * per].reshape(-1)

    # global references (one rank, no group)
    new_g = U.update_class_center_iter(feat, lab, cen, m=.9, num_class=k)
    sums_g = op.class_sums(feat, lab.reshape(-1), None, False, 0.0, None, 1, k)
    fg = feat.clone().requires_grad_(True)
    pg = probs.clone().requires_grad_(True)
    soft_g, _, _ = U.cal_centroid(fg, pg, pseudo_label=True, weighted_ave=True, partition=parts, n_class=k, part_id=part)
    (torch.cat(soft_g) * gcen).sum().backward()
    hard_g, _, _ = U.cal_centroid(feat, lab, n_class=k)

    def warm(r):                                        # allocator warm-up on each rank's stream (no exchange)
        U.update_class_center_iter(sl(feat, r), sl(lab, r), cen, m=.9, num_class=k)
        U.cal_centroid(sl(feat, r), sl(probs, r), pseudo_label=True, weighted_ave=True, partition=parts, n_class=k,
                       part_id=pix(part, r))
    _on_streams(streams, warm)
    torch.cuda.synchronize()

    res = {}

    def rank_fn(r):
        new = U.update_class_center_iter(sl(feat, r), sl(lab, r), cen, m=.9, num_class=k, group=boxes[r])
        _, sums = op.class_centres_update(sl(feat, r).contiguous(), sl(lab, r).reshape(-1), cen, .9, *boxes[r].args())
        hard, _, _ = U.cal_centroid(sl(feat, r), sl(lab, r), n_class=k, group=boxes[r])
        f = sl(feat, r).clone().requires_grad_(True)
        p = sl(probs, r).clone().requires_grad_(True)
        soft, _, _ = U.cal_centroid(f, p, pseudo_label=True, weighted_ave=True, partition=parts, n_class=k,
                                    part_id=pix(part, r), group=boxes[r])
        (torch.cat(soft) * gcen).sum().backward()
        res[r] = (new, sums, hard, torch.cat(soft).detach(), f.grad, p.grad)
    _on_streams(streams, rank_fn)
    torch.cuda.synchronize()
    assert all(bx.timeouts() == 0 for bx in boxes)
    for r in range(world):
        new, sums, hard, soft, df, dp = res[r]
        assert torch.equal(sums[:, -1], sums_g[:, -1])                        # class counts: bit-exact, GLOBAL
        assert float(sums[3, -1]) == 0.0 and torch.equal(new[3], new_g[3])     # empty class keeps its old centre blend
        close(sums, sums_g, rtol=1e-6, atol=1e-6)
        close(new, new_g, rtol=1e-5)
        close(hard, hard_g, rtol=1e-5)
        close(soft, torch.cat(soft_g), rtol=1e-5)
        grad_close(df, sl(fg.grad, r), rtol=1e-5)
        grad_close(dp, sl(pg.grad, r), rtol=1e-5)
        assert torch.equal(res[r][0], res[0][0]) and torch.equal(res[r][3], res[0][3])     # identical bits on every rank


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_prototype_loss_and_target_step_equal_global(api, world):
    """mpcl_loss_calc(group=mailbox) / mpcl_target_step(group=mailbox): the loss is the mean over the GLOBAL batch and the
    gradients are the global gradient's shard (utils/loss.py:558-571 over all ranks' pixels)."""
    from slcl.peer import LoopbackMailboxes
    L, _ = api
    b, c, h, w, k = 2 * world, 64, 20, 28, 5
    g = cases.g(41)
    feat = torch.randn(b, c, h, w, generator=g).to(dev())
    lab = torch.randint(0, k, (b, h, w), generator=g).to(dev())
    sel = (torch.rand(b * h * w, generator=g) > 0.4).float().to(dev())
    cen = torch.randn(k, c, generator=g).to(dev())
    mp = L.MPCL(dev(), num_class=k, temperature=.1, base_temperature=1, m=.4)
    per = b // world
    sl = lambda t, r: t[r * per:(r + 1) * per]
    pix = lambda t, r: t.reshape(b, -1)[r * per:(r + 1) * per].reshape(-1)

    def three(f, labels, selv, group):
        l1 = L.mpcl_loss_calc(f, labels, cen, mp, tag='source', group=group)
        l2 = L.mpcl_loss_calc(f, labels.reshape(-1), cen, mp, pixel_sel_loc=selv, tag='target', group=group)
        l3, hard, m = L.mpcl_target_step(f, cen, mp, .05, group=group)
        return l1, l2, l3

    fg = feat.clone().requires_grad_(True)
    want = three(fg, lab, sel, None)
    (want[0] + 2 * want[1] + 3 * want[2]).backward()
    boxes = LoopbackMailboxes(world, dev()).boxes
    streams = _rank_streams(world)
    _on_streams(streams, lambda r: sum(three(sl(feat, r).clone().requires_grad_(True), sl(lab, r), pix(sel, r), None)).backward())
    torch.cuda.synchronize()
    res = {}

    def rank_fn(r):
        f = sl(feat, r).clone().requires_grad_(True)
        got = three(f, sl(lab, r), pix(sel, r), boxes[r])
        (got[0] + 2 * got[1] + 3 * got[2]).backward()
        res[r] = ([x.detach() for x in got], f.grad)
    _on_streams(streams, rank_fn)
    torch.cuda.synchronize()
    assert all(bx.timeouts() == 0 for bx in boxes)
    for r in range(world):
        for a_, b_ in zip(res[r][0], want):
            close(a_, b_, rtol=1e-6, atol=1e-9)
        grad_close(res[r][1], sl(fg.grad, r), rtol=1e-5)


@pytest.mark.parametrize("world,with_sel", [(2, True), (4, False)])
def test_split_phase_exchange_plan_equals_global(world, with_sel):
    """ProtoPlan.forward(mailbox, split_phase=True) + backward(mailbox): the finaliser only sends, the backward kernel
    receives (block 0) and publishes to its other blocks.  Ranks = streams of one GPU; result must equal the global batch,
    over several steps (epoch parities, ready/pending words)."""
    from slcl.peer import LoopbackMailboxes
    from slcl.plan import ProtoPlan
    b, c, h, w, k = 2 * world, 32, 32, 24, 4
    g = cases.g(61 + world)
    feat = torch.randn(b, c, h, w, generator=g).to(dev())
    lab = torch.randint(0, k, (b * h * w,), generator=g).to(dev())
    sel = (torch.rand(b * h * w, generator=g) > 0.4).float().to(dev()) if with_sel else None
    cen = torch.randn(k, c, generator=g).to(dev())
    whole = ProtoPlan(feat, lab, sel, cen, k, .1, 1.0, .4)
    whole.forward(); whole.backward()
    torch.cuda.synchronize()
    per = b // world
    hw = h * w
    plans = [ProtoPlan(feat[r * per:(r + 1) * per].contiguous(), lab[r * per * hw:(r + 1) * per * hw].contiguous(),
                       None if sel is None else sel[r * per * hw:(r + 1) * per * hw].contiguous(), cen, k, .1, 1.0, .4)
             for r in range(world)]
    boxes = LoopbackMailboxes(world, dev()).boxes
    streams = _rank_streams(world)
    for step in range(3):
        _on_streams(streams, lambda r: (plans[r].forward(boxes[r], split_phase=True), plans[r].backward(boxes[r])))
        torch.cuda.synchronize()
        for r in range(world):
            close(plans[r].scal[0], whole.scal[0], rtol=1e-6, atol=0)
            close(plans[r].scal[2], whole.scal[2], rtol=1e-6, atol=0)
            grad_close(plans[r].dfeat, whole.dfeat[r * per:(r + 1) * per], rtol=1e-5)
            assert torch.equal(plans[r].scal, plans[0].scal)
    assert all(bx.epoch() == 3 and bx.timeouts() == 0 for bx in boxes)


def test_c_abi_called_directly_with_ctypes():
    """The INTEGRATION.md stub: raw ctypes against include/slcl.h, no torch custom-op layer in between."""
    import ctypes as C
    from slcl._lib import LIB_PATH
    lib = C.CDLL(LIB_PATH)

    class MapT(C.Structure):
        _fields_ = [(n, C.c_int64) for n in ("batch", "channels", "pixels", "stride_b", "stride_c", "stride_p")]

    class ParamsT(C.Structure):
        _fields_ = [("n_class", C.c_int), ("temperature", C.c_float), ("base_temperature", C.c_float),
                    ("margin", C.c_float), ("easy_margin", C.c_int), ("normalize", C.c_int)]
    P = C.c_void_p
    lib.slcl_proto_workspace_bytes.restype = C.c_size_t
    lib.slcl_proto_workspace_bytes.argtypes = [C.c_int64]
    lib.slcl_proto_fwd.argtypes = [P, C.POINTER(MapT), P, P, P, P, C.POINTER(ParamsT), P, P, P, P, C.c_size_t, P]
    lib.slcl_proto_bwd.argtypes = [P, C.POINTER(MapT), P, P, P, P, C.POINTER(ParamsT), P, P]
    gen = cases.g(321)
    b, c, h, w, k = 2, 64, 16, 16, 5
    feat_h = torch.randn(b, c, h, w, generator=gen)
    lab_h = torch.randint(0, k, (b * h * w,), generator=gen)
    cen_h = torch.randn(k, c, generator=gen)
    feat, labels, centres = feat_h.to(dev()), lab_h.to(dev()), cen_h.to(dev())
    n = b * h * w
    m = MapT(b, c, h * w, c * h * w, h * w, 1)
    p = ParamsT(k, 0.1, 1.0, 0.4, 0, 1)
    stash = torch.empty(k + 1, n, device=dev())
    cstate = torch.empty(k * c + k, device=dev())
    scal = torch.empty(4, device=dev())
    ws = torch.empty(lib.slcl_proto_workspace_bytes(n), dtype=torch.uint8, device=dev())
    dfeat = torch.empty_like(feat)
    g = torch.ones(1, device=dev())
    s = torch.cuda.current_stream().cuda_stream
    assert lib.slcl_proto_fwd(feat.data_ptr(), C.byref(m), labels.data_ptr(), None, None, centres.data_ptr(), C.byref(p),
                              stash.data_ptr(), cstate.data_ptr(), scal.data_ptr(), ws.data_ptr(), ws.numel(), s) == 0
    assert lib.slcl_proto_bwd(feat.data_ptr(), C.byref(m), stash.data_ptr(), cstate.data_ptr(), scal.data_ptr(),
                              g.data_ptr(), C.byref(p), dfeat.data_ptr(), s) == 0
    # error path: too small a workspace is refused before any launch
    assert lib.slcl_proto_fwd(feat.data_ptr(), C.byref(m), labels.data_ptr(), None, None, centres.data_ptr(), C.byref(p),
                              stash.data_ptr(), cstate.data_ptr(), scal.data_ptr(), ws.data_ptr(), 8, s) == -3
    fo = feat_h.clone().requires_grad_(True)
    ref = O.mpcl_loss_calc(fo, lab_h.view(b, h, w), cen_h, O.MarginSpec(k, .1, .4, 1.0))
    ref.backward()
    close(scal[0], ref)
    grad_close(dfeat, fo.grad)


def test_host_buffer_pipeline_matches_device_path(api):
    """slcl.host: chunked H2D -> fwd -> bwd -> D2H pipeline == mpcl_loss_calc(...).backward() on the device."""
    loss_mod, _ = api
    from slcl.host import mpcl_loss_and_grad_host
    gen = cases.g(404)
    b, c, h, w, k = 8, 32, 16, 16, 4
    feat_h = torch.randn(b, c, h, w, generator=gen).pin_memory()
    lab_h = torch.randint(0, k, (b, h, w), generator=gen).pin_memory()
    sel_h = (torch.rand(b * h * w, generator=gen) > 0.4).float().pin_memory()
    cc = torch.randn(k, c, generator=gen).to(dev())
    mp = loss_mod.MPCL(dev(), num_class=k, temperature=.1, base_temperature=1, m=.2)
    for sel in (sel_h, None):
        loss, grad_h = mpcl_loss_and_grad_host(feat_h, lab_h, cc, mp, pixel_sel_loc_h=sel, device=dev(), chunk_images=2)
        torch.cuda.synchronize()
        f = feat_h.to(dev()).requires_grad_(True)
        ref = loss_mod.mpcl_loss_calc(f, lab_h.view(-1).to(dev()), cc, mp, pixel_sel_loc=None if sel is None else sel.to(dev()),
                                      tag='target')
        ref.backward()
        close(loss, ref, rtol=1e-5)
        grad_close(grad_h, f.grad.cpu(), rtol=1e-5)


def test_fused_target_step_equals_two_call_sequence(api, golden):
    """f-1: generate_pseudo_label + target mpcl_loss_calc in one read of the map == the reference call sequence."""
    loss_mod, utils_mod = api
    ft = cases.kat2().to(dev())
    cc = cases.shipped_centres().to(dev())
    mp = loss_mod.MPCL(dev(), num_class=4, temperature=.1, base_temperature=1, m=.2)
    f = ft.clone().requires_grad_(True)
    out, hard, sel = loss_mod.mpcl_target_step(f, cc, mp, .25)
    out.backward()
    assert np.array_equal(hard.cpu().numpy(), golden["kat2_label"]) and np.array_equal(sel.cpu().numpy(), golden["kat2_sel"])
    close(out, golden["kat2_loss"], atol=1e-8)
    grad_close(f.grad, golden["kat2_dfeas"])
    # bit-identical to the unfused product path on a bigger, ragged map (scalar path) and a vector-path map
    for shape, k in (((3, 20, 7, 9), 5), ((2, 128, 16, 16), 5)):
        gen = cases.g(sum(shape))
        x = torch.randn(*shape, generator=gen).to(dev())
        cen = torch.randn(k, shape[1], generator=gen).to(dev())
        mpk = loss_mod.MPCL(dev(), num_class=k, temperature=.1, base_temperature=1, m=.2)
        h2, s2 = utils_mod.generate_pseudo_label(x, cen, .05)
        x1 = x.clone().requires_grad_(True)
        l2 = loss_mod.mpcl_loss_calc(x1, h2, cen, mpk, pixel_sel_loc=s2, tag='target')
        l2.backward()
        x2 = x.clone().requires_grad_(True)
        l1, h1, s1 = loss_mod.mpcl_target_step(x2, cen, mpk, .05)
        l1.backward()
        assert torch.equal(h1, h2) and torch.equal(s1, s2)
        assert torch.equal(l1, l2) and torch.equal(x1.grad, x2.grad)


# ---------------------------------------------------------------------------
# f-2: segmentation losses + entropy map
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("b,c,h,w,k,by_sel", [(2, 32, 16, 16, 4, False), (3, 32, 24, 20, 4, True), (4, 128, 32, 32, 5, False),
                                                (2, 64, 12, 12, 8, True), (2, 20, 8, 6, 3, False), (2, 160, 16, 16, 4, False)])
def test_fused_target_step_with_centroids_equals_separate_calls(api, b, c, h, w, k, by_sel):
    """f-1: pseudo labels + target loss + hard target centroids in ONE pass over F_t (slcl_target_step) must equal
    generate_pseudo_label -> mpcl_loss_calc(target) -> cal_centroid(one-hot of those labels): labels / mask bit-exact, loss
    and centroids to fp32 round-off, class counts exact, gradients equal.  C = 160 exercises the fallback (two reads)."""
    loss_mod, utils_mod = api
    g = cases.g(b * 100 + c + k)
    feat = torch.randn(b, c, h, w, generator=g).to(dev())
    cen = torch.randn(k, c, generator=g).to(dev())
    prev = torch.randn(k, c, generator=g).to(dev())
    gcen = torch.randn(k, c, generator=g).to(dev())
    mp = loss_mod.MPCL(dev(), num_class=k, temperature=.1, base_temperature=1, m=.2)
    # separate calls
    fs = feat.clone().requires_grad_(True)
    hard_s, sel_s = utils_mod.generate_pseudo_label(fs, cen, .05)
    loss_s = loss_mod.mpcl_loss_calc(fs, hard_s, cen, mp, pixel_sel_loc=sel_s, tag='target')
    lab_w = torch.where(sel_s > 0, hard_s, torch.full_like(hard_s, -1)) if by_sel else hard_s
    cen_s, _, _ = utils_mod.cal_centroid(fs, lab_w.view(b, h, w), previous_centroid=prev, momentum=.9, n_class=k)
    (loss_s + (cen_s * gcen).sum()).backward()
    # fused
    ff = feat.clone().requires_grad_(True)
    loss_f, hard_f, sel_f, cen_f = loss_mod.mpcl_target_step(ff, cen, mp, .05, with_centroids=True, weight_by_sel=by_sel,
                                                              previous_centroid=prev, momentum=.9)
    (loss_f + (cen_f * gcen).sum()).backward()
    assert torch.equal(hard_f, hard_s) and torch.equal(sel_f, sel_s)
    close(loss_f, loss_s, rtol=1e-6, atol=0)
    close(cen_f, cen_s, rtol=1e-5, atol=1e-6)
    grad_close(ff.grad, fs.grad, rtol=1e-5)
    if (h * w) % 4 == 0 and c <= 128:
        from slcl import ops as slcl_ops
        assert slcl_ops.target_step_supported(feat, k)
        out = torch.ops.slcl.target_step(feat, cen, .05, by_sel, k, .1, 1.0, .2, False, None, .9)
        want = torch.bincount(lab_w[lab_w >= 0], minlength=k)
        assert torch.equal(out[5][:, -1].long(), want)                           # class counts exact
        ref = torch.ops.slcl.proto_fwd_target(feat, cen, .05, k, .1, 1.0, .2, False)
        assert torch.equal(out[1], ref[1]) and torch.equal(out[2], ref[2])         # stash / cstate bit-identical


def test_source_step_equals_separate_calls(api):
    loss_mod, utils_mod = api
    g = cases.g(515)
    feat = torch.randn(4, 32, 24, 24, generator=g).to(dev())
    lab = torch.randint(0, 4, (4, 24, 24), generator=g).to(dev())
    cen = torch.randn(4, 32, generator=g).to(dev())
    mp = loss_mod.MPCL(dev(), num_class=4, temperature=.1, base_temperature=1, m=.4)
    fs = feat.clone().requires_grad_(True)
    c1 = utils_mod.update_class_center_iter(fs, lab, cen, m=.9)
    l1 = loss_mod.mpcl_loss_calc(fs, lab, c1.detach(), mp, tag='source')
    l1.backward()
    ff = feat.clone().requires_grad_(True)
    c2, l2 = loss_mod.mpcl_source_step(ff, lab, cen, mp, m=.9, num_class=4)
    l2.backward()
    assert torch.equal(c1, c2) and torch.equal(l1, l2) and torch.equal(fs.grad, ff.grad)


def test_seg_losses_vs_reference_golden_gpu():
    import os
    from slcl import seg
    gold = dict(np.load(os.path.join(cases.GOLDEN_DIR, "reference_seg_outputs.npz")))
    logits, labels = cases.seg_case()
    lab = labels.to(dev())
    for name, fn in (("ce", lambda z: seg.loss_calc(z, lab, 0, False)),
                     ("ce_jac", lambda z: seg.loss_calc(z, lab, 0, True)),
                     ("jac", lambda z: seg.jaccard_loss(lab, z)),
                     ("dice", lambda z: seg.dice_loss(z, lab)),
                     ("mpscl_seg", lambda z: seg.loss_calc(z, lab, 0, False) + seg.dice_loss(z, lab))):
        z = logits.to(dev()).requires_grad_(True)
        val = fn(z)
        (val * 1.5).backward()
        close(val, gold[f"seg_{name}_loss"])
        grad_close(z.grad / 1.5, gold[f"seg_{name}_dlogits"])
    z = logits.to(dev()).requires_grad_(True)
    ent = seg.prob_2_entropy(torch.softmax(z, dim=1))
    (ent * torch.linspace(0.5, 1.5, ent.numel(), device=dev()).view_as(ent)).sum().backward()
    close(ent, gold["seg_entropy_map"], atol=1e-7)
    grad_close(z.grad, gold["seg_entropy_dlogits"])
    # all three losses from ONE pass
    z = logits.to(dev())
    all3 = seg.seg_losses(z, lab)
    close(all3, [float(gold["seg_ce_loss"]), float(gold["seg_dice_loss"]), float(gold["seg_jac_loss"])])


@pytest.mark.parametrize("b,k,h,w", [(2, 4, 224, 224), (3, 5, 33, 33), (1, 2, 7, 5)])
def test_seg_losses_vs_oracle_shapes(b, k, h, w):
    from slcl import seg
    gen = cases.g(b * 1000 + k * 100 + h)
    logits = 1.5 * torch.randn(b, k, h, w, generator=gen)
    labels = torch.randint(0, k, (b, h, w), generator=gen)
    zo = logits.clone().requires_grad_(True)
    ref = O.loss_calc(zo, labels, True) + 0.5 * O.dice_loss(zo, labels)
    ref.backward()
    z = logits.to(dev()).requires_grad_(True)
    out = seg.loss_calc(z, labels.to(dev()), 0, True) + 0.5 * seg.dice_loss(z, labels.to(dev()))
    out.backward()
    close(out, ref)
    grad_close(z.grad, zo.grad)


def test_cross_entropy_ignores_unlabelled_pixels_like_nn_cross_entropy():
    """nn.CrossEntropyLoss (utils/loss.py:63-64) averages over the pixels it does not ignore (ignore_index = -100):
    ignored pixels count neither in the numerator nor in the denominator and get no gradient."""
    from slcl import seg
    gen = cases.g(9090)
    logits = 1.5 * torch.randn(3, 4, 20, 24, generator=gen)
    labels = torch.randint(0, 4, (3, 20, 24), generator=gen)
    labels[torch.rand(3, 20, 24, generator=gen) < 0.3] = -100
    zo = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(zo, labels, ignore_index=-100)
    ref.backward()
    z = logits.to(dev()).requires_grad_(True)
    out = seg.loss_calc(z, labels.to(dev()), 0, False)
    out.backward()
    close(out, ref)
    grad_close(z.grad, zo.grad)
    assert float(z.grad.permute(0, 2, 3, 1)[(labels == -100).to(dev())].abs().max()) == 0.0


def test_iscl_vs_reference_golden_gpu():
    """f-4: ISCL = two sweeps of the tensor-core kernel (bf16 inputs: rtol 2e-2)."""
    import os
    from slcl import losses
    gold = dict(np.load(os.path.join(cases.GOLDEN_DIR, "reference_seg_outputs.npz")))
    feats, l1, l2, dom, lam = cases.iscl_case()
    f = feats.to(dev()).requires_grad_(True)
    val = losses.InterpolatedSupervisedContrastiveLoss(0.5)(f, l1.to(dev()), l2.to(dev()), dom.to(dev()), lam.to(dev()))
    val.backward()
    close(val, gold["iscl_loss"], rtol=P2P_RTOL)
    grad_close(f.grad, gold["iscl_dfeat"], rtol=P2P_RTOL, floor=0.5)


def test_stock_signatures_reach_the_analytic_sweeps_and_stay_safe(api):
    """SupConLoss() / LocalConLoss() exactly as the reference constructs them (no n_class): class-index labels take the
    analytic sweeps (same value as the general sweeps and the oracle); labels outside [0, 8) at the first call select the
    general sweeps; labels that LEAVE the range later poison the loss with NaN instead of returning a wrong value."""
    loss_mod, _ = api
    gen = cases.g(808)
    f5 = F.normalize(torch.randn(2, 2, 24, 16, 16, generator=gen), dim=2)
    lab = torch.randint(0, 4, (2, 2, 16, 16), generator=gen)
    ref = O.supcon_loss(f5, lab, 0.7)
    stock = loss_mod.SupConLoss(0.7)
    out = stock(f5.to(dev()), lab.to(dev()))
    assert stock._classes.auto == 8
    close(out, ref, rtol=P2P_RTOL)
    close(out, loss_mod.SupConLoss(0.7, n_class=0)(f5.to(dev()), lab.to(dev())), rtol=1e-3)
    close(loss_mod.LocalConLoss(0.7, 2)(f5.to(dev()), lab.to(dev())), O.local_con_loss(f5, lab, 0.7, 2), rtol=P2P_RTOL)
    wild = lab * 37 - 5                                   # arbitrary integers: equality structure unchanged, 0 stays background?
    wild[lab == 0] = 0
    fresh = loss_mod.SupConLoss(0.7)
    close(fresh(f5.to(dev()), wild.to(dev())), ref, rtol=P2P_RTOL)      # first call sees them -> general sweeps
    assert fresh._classes.auto == 0
    assert torch.isnan(stock(f5.to(dev()), wild.to(dev())))               # `stock` decided "analytic" earlier: guarded


@pytest.mark.parametrize("n_class", [None, 0, 4])
def test_supcon_small_temperature_self_term_stays_finite(api, n_class):
    """at T = 0.07 the self pair dominates a row's exp-sum; the analytic / self-map paths subtract it after the
    sweep.  With a moderate number of rows the result must still match the oracle, and in the degenerate case (every other
    row orthogonal or opposite: the true sum of the others underflows) it must stay finite instead of log(<= 0)."""
    loss_mod, _ = api
    gen = cases.g(1717)
    f5 = F.normalize(torch.randn(1, 2, 32, 8, 8, generator=gen), dim=2)
    lab = torch.randint(0, 4, (1, 2, 8, 8), generator=gen)
    ref = O.supcon_loss(f5, lab, 0.07)
    fd = f5.to(dev()).requires_grad_(True)
    out = loss_mod.SupConLoss(0.07, n_class=n_class)(fd, lab.to(dev()))
    out.backward()
    close(out, ref, rtol=P2P_RTOL)
    assert torch.isfinite(fd.grad).all()
    # degenerate: 8 rows = +-e_1 .. +-e_4 (orthogonal or opposite), labels pair up the opposites
    eye = torch.eye(4)
    rows = torch.cat([eye, -eye]).reshape(1, 2, 4, 4, 1).permute(0, 1, 3, 2, 4).contiguous()      # [b=1, v=2, c=4, h=4, w=1]
    labd = torch.tensor([1, 2, 3, 1, 1, 2, 3, 1]).reshape(1, 2, 4, 1)
    outd = loss_mod.SupConLoss(0.07, n_class=n_class)(rows.to(dev()), labd.to(dev()))
    assert torch.isfinite(outd)


def test_large_one_row_set_problems_take_the_sorted_path(api):
    """>= 8192 rows over one row set: slcl.p2p gathers the rows sorted by label and hands the general sweeps self maps
    (label-uniform tiles on the fast path, no id tests).  SupConLoss (labelled) and ISCL against the oracle."""
    loss_mod, _ = api
    from slcl import losses, p2p
    assert 2 * 64 * 64 >= p2p._SORT_MIN_ROWS
    gen = cases.g(4242)
    f5 = F.normalize(torch.randn(1, 2, 16, 64, 64, generator=gen), dim=2)
    lab = torch.randint(0, 5, (1, 2, 64, 64), generator=gen)
    fo = f5.clone().requires_grad_(True)
    ref = O.supcon_loss(fo, lab, 0.7)
    ref.backward()
    for n_class in (0, None):                    # 0: general sweeps on sorted rows; None: the stock signature (auto -> analytic)
        f = f5.to(dev()).requires_grad_(True)
        out = loss_mod.SupConLoss(0.7, n_class=n_class)(f, lab.to(dev()))
        out.backward()
        close(out, ref, rtol=P2P_RTOL)
        grad_close(f.grad, fo.grad, rtol=P2P_RTOL, floor=0.5)
    n, d = 8192, 48
    feats = torch.randn(n, d, generator=gen)
    l1 = torch.randint(0, 6, (n,), generator=gen)
    l2 = torch.randint(0, 6, (n,), generator=gen)
    lam = torch.rand(n, generator=gen)
    dom = torch.where(lam > 0.5, l1, l2)
    fo = feats.clone().requires_grad_(True)
    ref = O.iscl_loss(fo, l1, l2, dom, lam, 0.5)
    ref.backward()
    f = feats.to(dev()).requires_grad_(True)
    val = losses.InterpolatedSupervisedContrastiveLoss(0.5)(f, l1.to(dev()), l2.to(dev()), dom.to(dev()), lam.to(dev()))
    val.backward()
    close(val, ref, rtol=P2P_RTOL)
    grad_close(f.grad, fo.grad, rtol=P2P_RTOL, floor=0.5)


def test_first_cuda_activity_of_the_autograd_thread_can_be_ours():
    """Regression: the tensor-core backward builds TMA descriptors with a driver-API call; in a fresh process the
    autograd worker thread has no current context until a runtime call binds one (CUDA_ERROR_INVALID_CONTEXT otherwise).
    Run the sampled pixel<->pixel loss fwd+bwd and a centre-gradient backward as the first things a new process does."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import sys, torch
sys.path.insert(0, %r)
from slcl import p2p, loss
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(1)
feat = torch.randn(2, 40, 24, 24, generator=g).to(dev).requires_grad_(True)
labels = torch.randint(0, 4, (2, 24, 24), generator=g).to(dev)
a_idx, c_idx = p2p.sample_class_balanced(labels, 200, 600, 4, torch.Generator().manual_seed(5))
out = p2p.sampled_supcon_loss(feat, labels, 200, 600, 4, temperature=0.7, anchor_idx=a_idx, contrast_idx=c_idx)
out.backward()
assert torch.isfinite(feat.grad).all()
print('p2p ok', float(out))
""" % os.path.join(root, "soft-labeled-contrastive-learning_b200")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "p2p ok" in res.stdout, res.stderr[-2000:]
    code2 = """
import sys, torch
sys.path.insert(0, %r)
from slcl import loss
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(2)
feat = torch.randn(2, 64, 16, 16, generator=g).to(dev).requires_grad_(True)
labels = torch.randint(0, 5, (2, 16, 16), generator=g).to(dev)
cen = torch.randn(5, 64, generator=g).to(dev).requires_grad_(True)
mp = loss.MPCL(dev, num_class=5, temperature=.1, base_temperature=1, m=.4)
out = loss.mpcl_loss_calc(feat, labels, cen, mp)
out.backward()
assert torch.isfinite(cen.grad).all() and torch.isfinite(feat.grad).all()
print('proto ok', float(out))
""" % os.path.join(root, "soft-labeled-contrastive-learning_b200")
    res = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "proto ok" in res.stdout, res.stderr[-2000:]
