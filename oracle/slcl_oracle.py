"""Plain-PyTorch CPU restatement of the SLCL loss hot path (TEST INFRASTRUCTURE).

Every function restates one reference function of the path named by
BASELINE.json:north_star and cites the reference ``file:line`` it follows
(paths relative to the reference repo root).  The restatement keeps the
reference's *operation order* where fp32 rounding depends on it (so it can be
compared bit-tightly with the reference and timed as an honest CPU port), but
is written independently: tensors stay wherever the caller put them (no
``.cuda()`` calls), there are no host syncs, and K / partitions are generic.

Gradients come from torch autograd over these functions.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this file.  See oracle/__init__.py for the pinning status of
each function.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# a-1  MPCL.forward                                    utils/loss.py:469-573
# --------------------------------------------------------------------------
@dataclass
class MarginSpec:
    """Constructor state of the reference ``MPCL`` (utils/loss.py:470-482)."""
    num_class: int = 5
    temperature: float = 0.07
    m: float = 0.5
    base_temperature: float = 0.07
    easy_margin: bool = False

    @property
    def cos_m(self) -> float:
        return math.cos(self.m)

    @property
    def sin_m(self) -> float:
        return math.sin(self.m)

    @property
    def th(self) -> float:          # utils/loss.py:479
        return math.cos(math.pi - self.m)

    @property
    def mm(self) -> float:          # utils/loss.py:480
        return math.sin(math.pi - self.m) * self.m


def mpcl_forward(spec: MarginSpec, features: torch.Tensor, labels: Optional[torch.Tensor],
                 class_center_feas: torch.Tensor, pixel_sel_loc: Optional[torch.Tensor] = None,
                 mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pixel -> prototype margin InfoNCE.  Follows utils/loss.py:484-573.

    features [N,1,C] unit rows; labels [N] int or None; class_center_feas [C,K]
    unit columns; pixel_sel_loc [N] or None; mask [N,K] soft labels or None.
    """
    if features.dim() < 3:                                         # :494-496
        raise ValueError('`features` needs to be [bsz, n_views, ...],'
                         'at least 3 dimensions are required')
    if features.dim() > 3:                                         # :497-498
        features = features.reshape(features.shape[0], features.shape[1], -1)
    n = features.shape[0]
    if labels is not None and mask is not None:                    # :502-503
        raise ValueError('Cannot define both `labels` and `mask`')
    if labels is None and mask is None:                            # :504-505
        pos = torch.eye(n, dtype=torch.float32, device=features.device)
    elif labels is not None:                                       # :506-515
        lab = labels.contiguous().view(-1, 1).long()
        if lab.shape[0] != n:
            raise ValueError('Num of labels does not match num of features')
        classes = torch.arange(spec.num_class, device=lab.device).view(1, -1)
        pos = (lab == classes).float()
    else:                                                          # :516-517
        pos = mask.float()

    views = features.shape[1]                                      # :521-525
    anchors = torch.cat(torch.unbind(features, dim=1), dim=0)
    t = spec.temperature

    cosine = anchors @ class_center_feas                           # :529
    plain = cosine / t                                             # :530-532
    plain = plain - plain.max(dim=1, keepdim=True).values.detach()
    sine = torch.sqrt((1.0 - cosine.pow(2)).clamp(0.0001, 1.0))    # :534
    phi = cosine * spec.cos_m - sine * spec.sin_m                  # :536
    if spec.easy_margin:                                           # :538-541
        phi = torch.where(cosine > 0, phi, cosine)
    else:
        phi = torch.where(cosine > spec.th, phi, cosine - spec.mm)
    marg = phi / t                                                 # :543-546
    marg = marg - marg.max(dim=1, keepdim=True).values.detach()

    pos = pos.repeat(views, views)                                 # :548
    z = plain * (1 - pos) + marg * pos                             # :550-554
    log_prob = z - torch.log(torch.exp(z).sum(1, keepdim=True) + 1e-4)   # :556
    row = (pos * log_prob).sum(1)                                  # :562 / :568
    scale = spec.temperature / spec.base_temperature
    if pixel_sel_loc is not None:                                  # :558-565
        sel = pixel_sel_loc.view(-1)
        return (-scale * (sel * row)).sum() / (sel.sum() + 1e-4)
    return (-scale * row).view(views, n).mean()                    # :568-571


# --------------------------------------------------------------------------
# a-2  mpcl_loss_calc                                  utils/loss.py:576-605
# --------------------------------------------------------------------------
def nearest_label_resize(labels: torch.Tensor, fea_h: int, fea_w: int) -> torch.Tensor:
    """Two 1-D nearest interpolations, W first then H (utils/loss.py:585-590)."""
    lab = labels.float()
    lab = F.interpolate(lab, size=fea_w, mode='nearest')
    lab = lab.permute(0, 2, 1).contiguous()
    lab = F.interpolate(lab, size=fea_h, mode='nearest')
    return lab.permute(0, 2, 1).contiguous()


def mpcl_loss_calc(feas: torch.Tensor, labels: torch.Tensor, class_center_feas: torch.Tensor,
                   spec: MarginSpec, pixel_sel_loc: Optional[torch.Tensor] = None,
                   tag: str = 'source') -> torch.Tensor:
    """NCHW features + [K,C] centres -> scalar loss (utils/loss.py:576-605)."""
    n, c, fea_h, fea_w = feas.shape
    if tag == 'source' and (labels.shape[1] != fea_h or labels.shape[2] != fea_w):
        labels = nearest_label_resize(labels, fea_h, fea_w)
    labels = labels.reshape(-1).long()                             # :592-593
    unit = F.normalize(feas, p=2, dim=1)                           # :595
    unit = unit.transpose(1, 2).transpose(2, 3).contiguous()       # :596
    unit = unit.reshape(n * fea_h * fea_w, c).unsqueeze(1)         # :597-598
    centers = F.normalize(class_center_feas, p=2, dim=1).transpose(0, 1)   # :600-601
    return mpcl_forward(spec, unit, labels, centers, pixel_sel_loc=pixel_sel_loc)


# --------------------------------------------------------------------------
# a-3  update_class_center_iter                     utils/utils_.py:568-594
# --------------------------------------------------------------------------
def update_class_center_iter(cla_src_feas: torch.Tensor, batch_src_labels: torch.Tensor,
                             class_center_feas: torch.Tensor, m: float = .2,
                             num_class: int = 4) -> torch.Tensor:
    feats = cla_src_feas.detach()                                  # :573
    lab = batch_src_labels.unsqueeze(1)                            # :578
    rows = []
    for k in range(num_class):                                     # :580-590
        sel = (lab == k).float()
        total = (feats * sel).sum(dim=[0, 2, 3])
        count = sel.sum()
        if count == 0:
            rows.append(class_center_feas[k, :].detach().unsqueeze(0))
        else:
            rows.append((total / count).unsqueeze(0))
    batch_centers = torch.cat(rows, dim=0)
    return m * class_center_feas + (1 - m) * batch_centers         # :592


# --------------------------------------------------------------------------
# a-4  generate_pseudo_label                        utils/utils_.py:597-624
# --------------------------------------------------------------------------
def generate_pseudo_label(cla_feas_trg: torch.Tensor, class_centers: torch.Tensor,
                          pixel_sel_th: float = .25) -> Tuple[torch.Tensor, torch.Tensor]:
    f = cla_feas_trg.detach()                                      # :613
    _, c, _, _ = f.shape
    f = F.normalize(f, p=2, dim=1)                                 # :615
    cen = F.normalize(class_centers, p=2, dim=1)                   # :616
    f = f.transpose(1, 2).contiguous().transpose(2, 3).contiguous().reshape(-1, c)   # :617-618
    cosine = f @ cen.transpose(0, 1)                               # :619-620
    ordered, _ = torch.sort(cosine, dim=1)                         # :607-610
    gap = ordered[:, -1] - ordered[:, -2]
    sel = torch.where(gap > pixel_sel_th, torch.ones(1, device=f.device), torch.zeros(1, device=f.device))
    return torch.argmax(cosine, dim=1), sel                        # :621-624


# --------------------------------------------------------------------------
# a-5  cal_centroid (repaired) + rMC partitions     utils/utils_.py:479-565
# --------------------------------------------------------------------------
def rmc_partition_ids(n_pixels: int, partition: int, generator: torch.Generator,
                      device: Optional[torch.device] = None) -> torch.Tensor:
    """Reversed-Monte-Carlo partition assignment.  PARITY UNPINNED: the reference
    accepts ``partition`` (utils/utils_.py:479) but never implements the
    sampler; this is the spec of SURVEY.md section 8(c)-2: a random permutation
    of the pixel indices (NCHW pixel order b,h,w) taken modulo P, so partitions
    are balanced to +-1 pixel."""
    perm = torch.randperm(n_pixels, generator=generator, device=generator.device)
    ids = (perm % partition).to(torch.int32)
    return ids if device is None else ids.to(device)


def cal_centroid(decoder_ft: torch.Tensor, label: torch.Tensor, previous_centroid=None,
                 momentum: float = 0.95, pseudo_label: bool = False, n_class: int = 4,
                 partition: int = 1, threshold=None, thd_w: float = 0.0, weighted_ave: bool = False,
                 epoch: int = 0, max_epoch: int = 1000, low_thd=0, high_thd=0.99, stdmin: bool = False,
                 part_id: Optional[torch.Tensor] = None):
    """Per-class (weighted) mean of decoder features.

    Follows utils/utils_.py:479-565 with the one-line repair of SURVEY.md
    section 0 (the inner list is created before the class loop at :516).  For
    ``partition > 1`` (not implemented in the reference) ``part_id`` [B*H*W]
    int, values in [0,P), restricts partition p to its own pixels and a list
    of P tensors is returned, as the caller expects
    (trainer/Trainer_MCCL.py:281-287).
    """
    b, c, h, w = decoder_ft.shape
    lab = label
    if not pseudo_label and (lab.shape[-1] != w or lab.shape[-2] != h):     # :498-502
        if lab.dim() == 3:
            lab = lab.unsqueeze(1)
        lab = F.interpolate(lab.float(), size=(h, w), mode='nearest').long().squeeze(1)
    if pseudo_label and (lab.shape[-1] != w or lab.shape[-2] != h):         # :503-505
        if lab.dim() == 3:
            raise ValueError("Soft pseudo-label must have channel dimension K")
        lab = F.interpolate(lab, size=(h, w), mode='bilinear', align_corners=False)

    if partition > 1:
        if part_id is None:
            raise ValueError("partition > 1 needs part_id (see rmc_partition_ids)")
        part_maps = [(part_id.view(b, 1, h, w) == p).to(decoder_ft.dtype) for p in range(partition)]
    else:
        part_maps = [None]

    sets: List[torch.Tensor] = []
    for pm in part_maps:
        rows = []
        if pseudo_label:                                                     # :509-530
            onehot = F.one_hot(torch.argmax(lab, dim=1), num_classes=n_class).permute(0, 3, 1, 2).float()
            certain = torch.ones_like(lab[:, 0:1])
            if threshold is not None and 0 < threshold < 1:
                certain = (lab.max(dim=1, keepdim=True).values >= threshold).float()
            if pm is not None:
                certain = certain * pm
            for k in range(n_class):
                src = lab if weighted_ave else onehot
                wk = src[:, k].unsqueeze(1) * certain
                total = (decoder_ft * wk).sum(dim=(0, 2, 3))
                rows.append((total / (wk.sum(dim=(0, 2, 3)).squeeze() + 1e-7)).unsqueeze(0))
            sets.append(torch.cat(rows, dim=0))
        else:                                                                # :532-540
            for k in range(n_class):
                wk = (lab == k).unsqueeze(1).float()
                if pm is not None:
                    wk = wk * pm
                total = (decoder_ft * wk).sum(dim=(0, 2, 3))
                rows.append(total / (wk.sum(dim=(0, 2, 3)).squeeze() + 1e-7))
            sets.append(torch.stack(rows, dim=0))

    if partition > 1:                                                        # :544-549
        out = sets
    else:
        out = sets[0]
    if previous_centroid is not None:                                        # :552-563
        if isinstance(out, list) and isinstance(previous_centroid, list) and len(out) == len(previous_centroid):
            out = [momentum * previous_centroid[i] + (1 - momentum) * out[i] for i in range(len(out))]
        elif isinstance(out, torch.Tensor) and isinstance(previous_centroid, torch.Tensor) \
                and out.shape == previous_centroid.shape:
            out = momentum * previous_centroid + (1 - momentum) * out
    return out, None, []                                                     # :565


# --------------------------------------------------------------------------
# a-6  ContrastiveLoss + CNR     utils/loss.py:233-275, Trainer_MCCL.py:303-315
# --------------------------------------------------------------------------
def contrastive_loss(centroid_s: torch.Tensor, centroid_t: torch.Tensor, bg: bool = False,
                     split: bool = False, norm: bool = True, n_rows: int = 4) -> torch.Tensor:
    """Centroid <-> centroid InfoNCE.  ``tau`` does not appear: the reference
    stores it and never uses it (utils/loss.py:236 vs :264-265).  ``n_rows=4``
    is the reference's hard-coded class count (:266)."""
    if norm:                                                       # :242-246
        centroid_t = centroid_t / (centroid_t.norm(p=2, dim=1, keepdim=True) + 1e-7)
        centroid_s = centroid_s / (centroid_s.norm(p=2, dim=1, keepdim=True) + 1e-7)
    cross = torch.exp(centroid_t @ centroid_s.t())                 # :264
    own = torch.exp(centroid_t @ centroid_t.t())                   # :265
    first = 0 if bg else 1
    idx = torch.arange(first, n_rows)
    den = cross[first:].sum(1) + own[first:].sum(1)                # :267
    if split:                                                      # :268-270
        rows = 0.5 * (-torch.log(cross[idx, idx] / (den + 1e-7)) - torch.log(own[idx, idx] / (den + 1e-7)))
    else:                                                          # :272-273
        rows = -torch.log((cross[idx, idx] + own[idx, idx]) / (den + 1e-7))
    return rows.sum()


def cnr_loss(centroid_s: torch.Tensor, centroid_t_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """Centroid-norm regulariser, trainer/Trainer_MCCL.py:303-315 (MSE of row norms / P)."""
    s_norm = centroid_s.norm(p=2, dim=1)
    total = 0
    for ct in centroid_t_list:
        total = total + F.mse_loss(ct.norm(p=2, dim=1), s_norm) / len(centroid_t_list)
    return total


# --------------------------------------------------------------------------
# a-7  SupConLoss                                      utils/loss.py:315-387
# --------------------------------------------------------------------------
def supcon_loss(features: torch.Tensor, labels: Optional[torch.Tensor] = None,
                temperature: float = 0.07) -> torch.Tensor:
    """All-pairs pixel <-> pixel supervised contrastive loss on [b,v,c,h,w]."""
    if features.dim() <= 3:                                        # :334-336
        raise ValueError('`features` needs to be [bsz, n_views, ...],'
                         'at least 4 dimensions are required')
    views = features.shape[1]
    stacked = torch.cat(torch.unbind(features, dim=1), dim=0)      # :337-340  [b*v,c,h,w]
    rows = stacked.permute(0, 2, 3, 1).reshape(-1, stacked.shape[1])   # :342-343  [M,c]
    gram = F.conv2d(stacked, rows.reshape(-1, stacked.shape[1], 1, 1)) / temperature   # :345-346
    gram = gram.permute(1, 0, 2, 3).reshape(gram.shape[1], -1)     # :347-349  [M,M]
    m = gram.shape[0]
    if labels is not None:                                         # :351-358
        lab = torch.cat(torch.unbind(labels, dim=1), dim=0).contiguous().view(-1, 1)
        pos = (lab == lab.t()).float()
        fg = (lab.squeeze() != 0).int()
    else:                                                          # :359-361
        pos = torch.eye(m // views, device=gram.device).repeat(views, views)
    off_diag = 1.0 - torch.eye(m, device=gram.device)              # :365-367
    pos = pos * off_diag                                           # :368
    e = torch.exp(gram) * off_diag                                 # :371
    log_prob = gram - torch.log(e.sum(1, keepdim=True))            # :372-373
    row = -(pos * log_prob).sum(1) / pos.sum(1)                    # :376-380
    if labels is not None:                                         # :382-386
        return (row * fg).sum() / fg.sum()
    return row.mean()


# --------------------------------------------------------------------------
# a-8  LocalConLoss / BlockConLoss                     utils/loss.py:390-466
# --------------------------------------------------------------------------
def local_con_loss(features: torch.Tensor, labels: Optional[torch.Tensor] = None,
                   temperature: float = 0.7, stride: int = 4) -> torch.Tensor:
    f = features[:, :, :, ::stride, ::stride]                      # :401
    if labels is None:
        return supcon_loss(f, None, temperature)
    lab = labels[:, :, ::stride, ::stride]                         # :404
    if lab.sum() == 0:                                             # :405-407
        return torch.zeros((), dtype=torch.float32, device=features.device)
    return supcon_loss(f, lab, temperature)


def block_con_loss(features: torch.Tensor, labels: Optional[torch.Tensor] = None,
                   temperature: float = 0.7, block_size: int = 32) -> torch.Tensor:
    div = features.shape[-1] // block_size                         # :426-428
    parts = []
    for i in range(div):                                           # :430-466
        for j in range(div):
            sl_h = slice(i * block_size, (i + 1) * block_size)
            sl_w = slice(j * block_size, (j + 1) * block_size)
            bf = features[:, :, :, sl_h, sl_w]
            if labels is not None:
                bl = labels[:, :, sl_h, sl_w]
                if bl.sum() == 0:                                  # :439-440
                    continue
                parts.append(supcon_loss(bf, bl, temperature))
            else:
                parts.append(supcon_loss(bf, None, temperature))
    if len(parts) == 0:                                            # :445-447
        return torch.zeros((), dtype=torch.float32, device=features.device)
    return torch.stack(parts).mean()


# --------------------------------------------------------------------------
# sampled rectangular pixel <-> pixel loss (cfg3).  PARITY UNPINNED (our spec,
# SURVEY.md section 8(c)-3); the square case A == B == all pixels is SupConLoss.
# --------------------------------------------------------------------------
def sample_class_balanced(labels_flat: torch.Tensor, n_pick: int, n_class: int, perm: torch.Tensor) -> torch.Tensor:
    """Class-balanced pick of exactly ``n_class * ceil(n_pick / n_class)`` distinct pixels from ONE random permutation
    ``perm`` of all pixel indices (``torch.randperm(N, generator)``, PyTorch's RNG stream):

      phase 1  per class k (labels in [0, n_class)): the first ``per = ceil(n_pick / n_class)`` pixels of class k in
               permutation order; output is class-major (class 0's picks, then class 1's, ...), each in permutation order;
      phase 2  when classes have fewer than ``per`` members, the open slots are filled -- after all phase-1 picks --
               with the first not-yet-picked valid pixels in permutation order, whatever their class.

    The output size does not depend on the class counts, so the CUDA path needs no size on the host.  With fewer than
    ``n_class * per`` valid pixels the tail repeats nothing: the function raises (the CUDA path returns the count and
    poisons the loss)."""
    lab = labels_flat.reshape(-1)[perm]
    per = -(-n_pick // n_class)
    total = per * n_class
    picked = torch.zeros(perm.numel(), dtype=torch.bool)
    out = []
    for k in range(n_class):
        pos = torch.nonzero(lab == k).squeeze(1)[:per]
        picked[pos] = True
        out.append(perm[pos])
    n1 = sum(o.numel() for o in out)
    valid = (lab >= 0) & (lab < n_class)
    rest = torch.nonzero(valid & ~picked).squeeze(1)[:total - n1]
    if n1 + rest.numel() < total:
        raise ValueError("not enough labelled pixels for the requested sample")
    out.append(perm[rest])
    return torch.cat(out)


def supcon_rect(anchor: torch.Tensor, contrast: torch.Tensor, anchor_lab: torch.Tensor,
                contrast_lab: torch.Tensor, anchor_idx: torch.Tensor, contrast_idx: torch.Tensor,
                temperature: float) -> torch.Tensor:
    """Rectangular SupCon: anchors [A,d] vs contrast rows [M,d].  Self pairs are
    those with equal *global pixel index*; positives have equal labels; row loss
    and foreground weighting as utils/loss.py:371-386."""
    s = (anchor @ contrast.t()) / temperature
    not_self = (anchor_idx.view(-1, 1) != contrast_idx.view(1, -1)).to(s.dtype)
    pos = (anchor_lab.view(-1, 1) == contrast_lab.view(1, -1)).to(s.dtype) * not_self
    e = torch.exp(s) * not_self
    log_prob = s - torch.log(e.sum(1, keepdim=True))
    row = -(pos * log_prob).sum(1) / pos.sum(1)
    fg = (anchor_lab != 0).to(s.dtype)
    return (row * fg).sum() / fg.sum()


def gather_unit_rows(feat_nchw: torch.Tensor, pixel_idx: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """Rows ``pixel_idx`` (b,h,w order) of an NCHW map, L2-normalised over C
    (eps 1e-12 like F.normalize), optionally rounded to bf16."""
    b, c, h, w = feat_nchw.shape
    rows = feat_nchw.permute(0, 2, 3, 1).reshape(-1, c)[pixel_idx]
    return F.normalize(rows, p=2, dim=1).to(dtype)


# --------------------------------------------------------------------------
# f-2  segmentation losses                     utils/loss.py:11-103, utils/utils_.py:627-631
# --------------------------------------------------------------------------
def jaccard_loss(true: torch.Tensor, logits: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """utils/loss.py:11-43 (num_classes > 1 branch)."""
    k = logits.shape[1]
    onehot = F.one_hot(true.reshape(true.shape[0], *logits.shape[2:]).long(), k).movedim(-1, 1).to(logits.dtype)   # :34-35
    probas = F.softmax(logits, dim=1)                              # :36
    dims = (0, 2, 3)
    inter = (probas * onehot).sum(dims)                            # :39
    card = (probas + onehot).sum(dims)                             # :40
    return 1 - (inter / (card - inter + eps)).mean()               # :41-43


def loss_calc(pred: torch.Tensor, label: torch.Tensor, jaccard: bool = False) -> torch.Tensor:
    """utils/loss.py:46-66."""
    loss = F.cross_entropy(pred, label.long())
    return loss + jaccard_loss(label, pred) if jaccard else loss


def dice_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """utils/loss.py:69-103."""
    n, c, h, w = pred.shape
    onehot = torch.zeros(n, c, h, w, dtype=pred.dtype).scatter_(1, target.unsqueeze(1).long(), 1)    # :77-79
    probs = F.softmax(pred, dim=1)                                 # :88
    num = (probs * onehot).sum(3).sum(2)                           # :89-91
    den1 = (probs * probs).sum(3).sum(2)                           # :93-95
    den2 = (onehot * onehot).sum(3).sum(2)                         # :97-99
    dice = 2.0 * (num / (den1 + den2 + 1e-5))                      # :101
    return 1 - dice.sum() / dice.size(0) / c                       # :103-104


def prob_2_entropy(prob: torch.Tensor) -> torch.Tensor:
    """utils/utils_.py:627-631."""
    return -prob * torch.log2(prob + 1e-7) / math.log2(prob.shape[1])


# --------------------------------------------------------------------------
# f-4  InterpolatedSupervisedContrastiveLoss           utils/losses.py:6-81
# --------------------------------------------------------------------------
def iscl_loss(features, labels_1, labels_2, dominant_labels, lambdas, temperature, normalize=True):
    if normalize:                                                  # :44-45
        features = F.normalize(features, dim=-1, p=2)
    s = features @ features.t() / temperature                      # :47-48
    s = s - s.max(dim=1, keepdim=True).values.detach()             # :50-51
    n = s.shape[0]
    off = ~torch.eye(n, dtype=torch.bool, device=s.device)         # :67-68

    def per_sample(query):                                         # :61-81
        pos = (query.unsqueeze(1) == dominant_labels.unsqueeze(0)) & off
        cnt = pos.sum(-1).to(s.dtype)
        return -(1 / cnt) * (pos * (s - torch.log((off * torch.exp(s)).sum(-1, keepdim=True)))).sum(-1)

    return (lambdas * per_sample(labels_1) + (1 - lambdas) * per_sample(labels_2)).mean()      # :53-59
