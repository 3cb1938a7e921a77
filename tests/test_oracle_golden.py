"""CPU tests: pin the oracle restatement (oracle/slcl_oracle.py) against

  * the outputs of the reference's own functions stored in
    tests/golden/reference_outputs.npz (made by oracle/make_golden.py), and
  * the nine known-answer values of SURVEY.md section 8(c).

Tolerances: the restatement performs the same fp32 operation sequence, so
losses agree to ~1e-6 relative; gradients to 1e-5.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cases
from oracle import slcl_oracle as O

RT = dict(rtol=2e-5, atol=1e-7)


def close(a, b, **kw):
    tol = dict(RT)
    tol.update(kw)
    np.testing.assert_allclose(np.asarray(a.detach() if torch.is_tensor(a) else a, dtype=np.float64),
                               np.asarray(b, dtype=np.float64), **tol)


def spec4(m):
    return O.MarginSpec(num_class=4, temperature=.1, base_temperature=1, m=m)


def test_kat1_source_loss_and_grads(golden):
    feas, labels = cases.kat1()
    cc = cases.shipped_centres()
    f = feas.clone().requires_grad_(True)
    c = cc.clone().requires_grad_(True)
    loss = O.mpcl_loss_calc(f, labels, c, spec4(.4), tag='source')
    loss.backward()
    assert abs(loss.item() - 0.23589777946472168) < 1e-6          # SURVEY KAT-1
    assert abs(f.grad.abs().sum().item() - 0.831756591796875) < 1e-5
    close(loss, golden["kat1_loss"])
    close(f.grad, golden["kat1_dfeas"], atol=1e-9)
    close(c.grad, golden["kat1_dcentres"], atol=1e-8)


def test_kat2_pseudo_label_and_target_loss(golden):
    ft = cases.kat2()
    cc = cases.shipped_centres()
    hard, sel = O.generate_pseudo_label(ft, cc, .25)
    assert torch.bincount(hard, minlength=4).tolist() == [40, 28, 27, 33]   # SURVEY KAT-2
    assert sel.sum().item() == 22
    assert np.array_equal(hard.numpy(), golden["kat2_label"])
    assert np.array_equal(sel.numpy(), golden["kat2_sel"])
    f = ft.clone().requires_grad_(True)
    loss = O.mpcl_loss_calc(f, hard, cc, spec4(.2), pixel_sel_loc=sel, tag='target')
    loss.backward()
    assert abs(loss.item() - 0.007330850698053837) < 1e-7
    close(f.grad, golden["kat2_dfeas"], atol=1e-9)


def test_kat3_ema_centres(golden):
    feas, labels = cases.kat1()
    cc = cases.shipped_centres()
    new = O.update_class_center_iter(feas, labels, cc, m=.9)
    assert abs(new.sum().item() - 4.65817403793335) < 1e-5          # SURVEY KAT-3
    close(new, golden["kat3_centres"])
    lab = labels.clone()
    lab[lab == 2] = 1
    new = O.update_class_center_iter(feas, lab, cc, m=.9)
    close(new, golden["kat3_empty_centres"])
    close(new[2], cc[2], rtol=1e-6)                                 # empty class keeps its centre (m*c+(1-m)*c)


def test_kat4_contrastive(golden):
    cs, ct = cases.kat4()
    for name, kw in (("plain", {}), ("split", {"split": True}), ("bg", {"bg": True})):
        a = cs.clone().requires_grad_(True)
        b = ct.clone().requires_grad_(True)
        loss = O.contrastive_loss(a, b, **kw)
        loss.backward()
        close(loss, golden[f"kat4_{name}_loss"])
        close(a.grad, golden[f"kat4_{name}_ds"], atol=1e-8)
        close(b.grad, golden[f"kat4_{name}_dt"], atol=1e-8)
    assert abs(O.contrastive_loss(cs, ct).item() - 2.896496295928955) < 1e-6
    close(golden["kat4_tau01_loss"], golden["kat4_plain_loss"], rtol=0, atol=0)   # tau ignored by the reference


def test_kat5_hard_centroids(golden):
    feas, labels = cases.kat1()
    ft = cases.kat2()
    c1, ratio, std = O.cal_centroid(feas, labels, momentum=.9)
    assert ratio is None and std == []
    f = ft.clone().requires_grad_(True)
    c2, _, _ = O.cal_centroid(f, labels, previous_centroid=c1.detach(), momentum=.9)
    (c2 * c2).sum().backward()
    assert abs(c1.sum().item() - 1.4735791683197021) < 1e-5          # SURVEY KAT-5
    assert abs(c2.sum().item() - 1.1630873680114746) < 1e-5
    close(c1, golden["kat5_c1"], atol=1e-8)
    close(c2, golden["kat5_c2"], atol=1e-8)
    close(f.grad, golden["kat5_dft"], atol=1e-9)


@pytest.mark.parametrize("name,kw", [
    ("wtd", dict(weighted_ave=True)),
    ("wtd_thd", dict(weighted_ave=True, threshold=0.6)),
    ("hardpl_thd", dict(weighted_ave=False, threshold=0.6)),
    ("wtd_ema", dict(weighted_ave=True, momentum=.9)),
])
def test_soft_centroids_vs_repaired_reference(golden, name, kw):
    sft, probs = cases.soft_case()
    if name == "wtd_ema":
        kw = dict(kw, previous_centroid=cases.shipped_centres())
    f = sft.clone().requires_grad_(True)
    p = probs.clone().requires_grad_(True)
    cen, _, _ = O.cal_centroid(f, p, pseudo_label=True, **kw)
    src_c, _ = cases.kat4()
    loss = O.contrastive_loss(src_c, cen) + (cen * cen).sum()
    loss.backward()
    close(cen, golden[f"soft_{name}_cen"], atol=1e-8)
    close(loss, golden[f"soft_{name}_loss"])
    close(f.grad, golden[f"soft_{name}_dft"], atol=1e-8)
    if golden[f"soft_{name}_dp"].size:
        close(p.grad, golden[f"soft_{name}_dp"], atol=1e-7)
    else:
        assert p.grad is None or float(p.grad.abs().sum()) == 0.0


def test_kat6_kat7_pixel_to_pixel(golden):
    f5, lab = cases.kat6()
    f = f5.clone().requires_grad_(True)
    loss = O.supcon_loss(f, lab, .7)
    loss.backward()
    assert abs(loss.item() - 4.883881568908691) < 2e-6                # SURVEY KAT-6
    close(loss, golden["kat6_loss"])
    close(f.grad, golden["kat6_dfeat"], atol=1e-8)
    f = f5.clone().requires_grad_(True)
    loss = O.supcon_loss(f, None, .7)
    loss.backward()
    close(loss, golden["kat6_unlab_loss"])
    close(f.grad, golden["kat6_unlab_dfeat"], atol=1e-8)
    f7, lab7 = cases.kat7()
    close(O.local_con_loss(f7, lab7, .7, 4), golden["kat7_local"])
    close(O.block_con_loss(f7, lab7, .7, 32), golden["kat7_block"])
    close(O.local_con_loss(f7, None, .7, 4), golden["kat7_local_unlab"])
    assert abs(float(golden["kat7_local"]) - 6.2668328285217285) < 1e-6   # SURVEY KAT-7
    assert abs(float(golden["kat7_block"]) - 7.655858516693115) < 1e-6


def test_rect_supcon_square_case_equals_supcon():
    """The rectangular loss (our spec) with A == B == all pixels is SupConLoss."""
    f5, lab = cases.kat6()
    stacked = torch.cat(torch.unbind(f5, dim=1), dim=0)
    rows = stacked.permute(0, 2, 3, 1).reshape(-1, 32)
    labs = torch.cat(torch.unbind(lab, dim=1), dim=0).reshape(-1)
    idx = torch.arange(rows.shape[0])
    a = O.supcon_rect(rows, rows, labs, labs, idx, idx, .7)
    close(a, O.supcon_loss(f5, lab, .7), rtol=1e-5)


def test_kat8_soft_mask(golden):
    feas, _ = cases.kat1()
    cc = cases.shipped_centres()
    unit = F.normalize(feas, p=2, dim=1).permute(0, 2, 3, 1).reshape(-1, 32).clone().requires_grad_(True)
    cen = F.normalize(cc, p=2, dim=1).t()
    loss = O.mpcl_forward(spec4(.4), unit.unsqueeze(1), None, cen, mask=cases.kat8_mask())
    loss.backward()
    assert abs(loss.item() - 0.2541099488735199) < 1e-6               # SURVEY KAT-8
    close(unit.grad, golden["kat8_dunit"], atol=1e-9)


def test_kat9_label_downsample(golden):
    f9, lab9, cc9 = cases.kat9()
    f = f9.clone().requires_grad_(True)
    loss = O.mpcl_loss_calc(f, lab9, cc9, O.MarginSpec(5, .1, .4, 1.0), tag='source')
    loss.backward()
    assert abs(loss.item() - 0.18390698730945587) < 1e-6              # SURVEY KAT-9
    close(f.grad.abs().sum(), golden["kat9_dfeas_abs_sum"], rtol=1e-4)
    close(f.grad[:, :4, :3, :3], golden["kat9_dfeas_head"], atol=1e-10)


def test_ragged_out_of_range_and_selection(golden):
    rf, rl, rc, rsel = cases.ragged_case()
    f = rf.clone().requires_grad_(True)
    c = rc.clone().requires_grad_(True)
    loss = O.mpcl_loss_calc(f, rl.view(-1), c, O.MarginSpec(5, .07, .5, .07), pixel_sel_loc=rsel, tag='target')
    loss.backward()
    close(loss, golden["ragged_loss"])
    close(f.grad, golden["ragged_dfeas"], atol=1e-8)
    close(c.grad, golden["ragged_dcentres"], atol=1e-7)
    hard, sel = O.generate_pseudo_label(rf, rc, .1)
    assert np.array_equal(hard.numpy(), golden["ragged_label"])
    assert np.array_equal(sel.numpy(), golden["ragged_sel"])
    close(O.update_class_center_iter(rf, rl, rc, m=.8, num_class=5), golden["ragged_ema"])
    f = rf.clone().requires_grad_(True)
    loss = O.mpcl_loss_calc(f, rl, rc, O.MarginSpec(5, .1, .4, 1.0, easy_margin=True), tag='source')
    loss.backward()
    close(loss, golden["ragged_easy_loss"])
    close(f.grad, golden["ragged_easy_dfeas"], atol=1e-9)


def test_error_behaviour_matches_reference():
    sp = spec4(.4)
    with pytest.raises(ValueError):
        O.mpcl_forward(sp, torch.randn(4, 8), torch.zeros(4), torch.randn(8, 4))
    with pytest.raises(ValueError):
        O.mpcl_forward(sp, torch.randn(4, 1, 8), torch.zeros(4), torch.randn(8, 4), mask=torch.ones(4, 4))
    with pytest.raises(ValueError):
        O.mpcl_forward(sp, torch.randn(4, 1, 8), torch.zeros(5), torch.randn(8, 4))
    with pytest.raises(ValueError):
        O.supcon_loss(torch.randn(2, 3, 4))


def test_rmc_partitions_balanced_and_reproducible():
    ids = O.rmc_partition_ids(1001, 3, cases.g(5))
    again = O.rmc_partition_ids(1001, 3, cases.g(5))
    assert torch.equal(ids, again)
    counts = torch.bincount(ids.long(), minlength=3)
    assert counts.max() - counts.min() <= 1
    ft, probs = cases.soft_case()
    pid = O.rmc_partition_ids(ft.shape[0] * ft.shape[2] * ft.shape[3], 2, cases.g(6))
    parts, _, _ = O.cal_centroid(ft, probs, pseudo_label=True, weighted_ave=True, partition=2, part_id=pid)
    assert isinstance(parts, list) and len(parts) == 2 and parts[0].shape == (4, 32)
    # partitions are a disjoint cover: weighted recombination gives the P=1 centroid
    whole, _, _ = O.cal_centroid(ft, probs, pseudo_label=True, weighted_ave=True)
    w = [(probs * (pid.view(2, 1, 12, 10) == p)).sum(dim=(0, 2, 3)) for p in range(2)]
    recombined = (parts[0] * w[0][:, None] + parts[1] * w[1][:, None]) / (w[0] + w[1])[:, None]
    close(recombined, whole, rtol=1e-4, atol=1e-6)


@pytest.fixture(scope="module")
def seg_golden():
    import os
    return dict(np.load(os.path.join(cases.GOLDEN_DIR, "reference_seg_outputs.npz")))


def test_seg_losses_vs_reference_golden(seg_golden):
    """f-2: loss_calc / jaccard_loss / dice_loss / prob_2_entropy restatements vs the reference's outputs."""
    logits, labels = cases.seg_case()
    for name, fn in (("ce", lambda z: O.loss_calc(z, labels, False)),
                     ("ce_jac", lambda z: O.loss_calc(z, labels, True)),
                     ("jac", lambda z: O.jaccard_loss(labels, z)),
                     ("dice", lambda z: O.dice_loss(z, labels)),
                     ("mpscl_seg", lambda z: O.loss_calc(z, labels, False) + O.dice_loss(z, labels))):
        z = logits.clone().requires_grad_(True)
        val = fn(z)
        val.backward()
        close(val, seg_golden[f"seg_{name}_loss"])
        close(z.grad, seg_golden[f"seg_{name}_dlogits"], atol=1e-8)
    z = logits.clone().requires_grad_(True)
    ent = O.prob_2_entropy(torch.softmax(z, dim=1))
    (ent * torch.linspace(0.5, 1.5, ent.numel()).view_as(ent)).sum().backward()
    close(ent, seg_golden["seg_entropy_map"], atol=1e-8)
    close(z.grad, seg_golden["seg_entropy_dlogits"], atol=1e-7)


def test_iscl_vs_reference_golden(seg_golden):
    feats, l1, l2, dom, lam = cases.iscl_case()
    f = feats.clone().requires_grad_(True)
    val = O.iscl_loss(f, l1, l2, dom, lam, 0.5)
    val.backward()
    close(val, seg_golden["iscl_loss"])
    close(f.grad, seg_golden["iscl_dfeat"], atol=1e-8)
