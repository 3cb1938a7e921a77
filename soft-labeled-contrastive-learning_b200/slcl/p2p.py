"""Pixel <-> pixel supervised contrastive losses on the tensor-core path.

Drop-ins for the reference's ``SupConLoss`` / ``LocalConLoss`` / ``BlockConLoss``
(utils/loss.py:315-466, duplicated at utils/losses.py:95-239) plus the sampled
rectangular variant of SURVEY.md 8(c)-3 (``sampled_supcon_loss``).  Rows are
gathered from the NCHW map into bf16 K-major rows, similarities run as
tcgen05 GEMMs with fp32 accumulation and a fused soft-max epilogue; the M x M
matrix of the reference (:342-349) is never materialised.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops  # noqa: F401

_ops = ops.dispatch          # eager: op bodies directly; compiled: torch.ops.slcl
_SORT_MIN_ROWS = 8192        # one-row-set problems at least this large are gathered sorted by label (see p2p_loss)
AUTO_N_CLASS = 8             # labelled SupCon family without an explicit n_class: analytic sweeps for labels in [0, 8)


class _AutoClasses:
    """Lets the STOCK signatures (``SupConLoss()``, ``LocalConLoss()``: no ``n_class`` argument) reach the analytic
    tensor-core sweeps.  Those need class-index labels in [0, n_class), n_class <= 8 -- true for every label map of the
    reference (4 or 5 classes, config.py:9) but not guaranteed by the signature.  The first labelled call of a module
    checks the range once on the host (one synchronisation in the module's lifetime).  Should a label leave the range
    in a later call the loss comes out NaN, never silently wrong, and at no cost: in these losses every row is an
    anchor, an anchor whose label is outside [0, n_class) has no class table, so the finishing kernel sees n_i = 0
    positives and its row loss is 0/0 -- which the weighted sum carries to the result whatever the row's weight.
    ``n_class=0`` in the constructor forces the general sweeps (any integer labels), ``n_class=k`` trusts the caller."""

    def __init__(self, n_class):
        self.fixed = n_class            # None = auto
        self.auto = None                # decided at the first labelled call

    def resolve(self, labels):
        """-> (n_class for the sweeps, device flag `labels in range` or None -- None since the kernels' own NaN covers it)"""
        if self.fixed is not None:
            return int(self.fixed), None
        if labels is None:
            return 0, None
        if self.auto is None:
            lo, hi = int(labels.min()), int(labels.max())        # once per module
            self.auto = AUTO_N_CLASS if (lo >= 0 and hi < AUTO_N_CLASS) else 0
        if self.auto == 0:
            return 0, None
        return self.auto, None


def _guard(loss, in_range):
    return loss if in_range is None else torch.where(in_range, loss, torch.full_like(loss, float("nan")))



class _P2PLoss(torch.autograd.Function):
    """loss(feat) for anchor rows idx_a and contrast rows idx_b of an NCHW map (appendix A.6).

    ``n_class`` > 0 selects the analytic sweeps (include/slcl.h): labels are class indices in [0, n_class), ids
    unique; the forward then also produces U and the per-class row sums, and the backward is one sweep."""

    @staticmethod
    def forward(ctx, feat, idx_a, idx_b, meta_a, meta_b, weight, temperature, normalize, same_rows, n_class, selfcol,
                selfrow, n_batch, rowmaps):
        fm = feat.detach().contiguous()
        c = fm.shape[1]
        b_bf16, _, inv_b = _ops.gather_unit_rows(fm, idx_b, normalize, True, False)
        if same_rows:
            a_bf16, inv_a = b_bf16, inv_b
        else:
            a_bf16, _, inv_a = _ops.gather_unit_rows(fm, idx_a, normalize, True, False)
        if normalize:
            shift = torch.full_like(inv_a, 1.0 / temperature)
        else:   # upper bound of S_ij: |a_i| max_j |b_j| / T
            shift = _ops.p2p_shift(inv_a, inv_b, temperature)
        keep = n_class > 0 and ctx.needs_input_grad[0]
        loss, stats, state = _ops.p2p_fwd(a_bf16, b_bf16, meta_a, meta_b, shift, weight, temperature, n_class, selfcol, keep,
                                          n_batch)
        ctx.save_for_backward(fm, idx_a, idx_b, meta_a, meta_b, weight, a_bf16, b_bf16, inv_a, inv_b, shift, stats, state,
                              selfcol, selfrow, rowmaps)
        ctx.cfg = (temperature, normalize, same_rows, c, n_class, n_batch)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        (fm, idx_a, idx_b, meta_a, meta_b, weight, a_bf16, b_bf16, inv_a, inv_b, shift, stats, state, selfcol,
         selfrow, rowmaps) = ctx.saved_tensors
        temperature, normalize, same_rows, c, n_class, n_batch = ctx.cfg
        if not ctx.needs_input_grad[0]:
            return (None,) * 14
        d_a, d_b = _ops.p2p_bwd(a_bf16, b_bf16, c, meta_a, meta_b, shift, weight, temperature, stats,
                                grad_out.reshape(1), True, True, n_class, selfcol, selfrow, state, n_batch)
        if rowmaps is not None and not same_rows:
            # pixel -> row maps at hand and the rows cover a good share of the map: write the whole gradient map from the
            # pixel side (coalesced, no atomics, no zero-fill) instead of scattering 4-byte elements
            dfeat = _ops.scatter_rows_by_map(fm, normalize, rowmaps[0], d_a.contiguous(), inv_a, rowmaps[1], d_b.contiguous(), inv_b)
            return (dfeat,) + (None,) * 13
        dfeat = torch.zeros_like(fm)
        if same_rows:       # anchors and contrast rows are the same gathered rows: one scatter of the summed row gradients
            _ops.scatter_rows_bwd(fm, idx_b, normalize, d_a + d_b, inv_b, dfeat)
        else:
            _ops.scatter_rows_bwd(fm, idx_a, normalize, d_a, inv_a, dfeat)
            _ops.scatter_rows_bwd(fm, idx_b, normalize, d_b, inv_b, dfeat)
        return (dfeat,) + (None,) * 13


def p2p_loss(feat, idx_a, idx_b, lab_a, lab_b, id_a, id_b, weight, temperature, normalize, same_rows=False, n_class=0,
             selfcol=None, selfrow=None, n_batch=1, metas=None, id_bound=None):
    """``same_rows``: anchors and contrast rows are the same gathered rows (one gather); their labels may
    still differ (ISCL compares query labels with the anchors' dominant labels).
    ``n_batch`` > 1: block-diagonal batches -- rows [z R/n, (z+1) R/n) of both sides only see each other (R/n a
    multiple of 128; either mode -- the analytic sweeps keep one table of per-class sums per batch).
    ``n_class`` in 1..8: labels are class indices in [0, n_class) and ids are unique -> analytic sweeps; the
    self-pair maps are derived from the ids unless given.
    ``metas``: the {label, id} rows (``ops.rows_meta`` / ``ops.pad_meta``) when the caller has built them already
    (``lab_*`` are then unused); ``id_bound``: the ids are int64 in [0, id_bound) (pixel indices) -> the self-pair maps
    come from two lookup tables instead of a sort."""
    if same_rows and n_class == 0 and n_batch == 1 and metas is None and idx_a.numel() >= _SORT_MIN_ROWS:
        # General labels over ONE large row set (SupCon family, ISCL): the sums are order-invariant, so gather the rows
        # sorted by contrast label -- label-uniform column tiles take the sweeps' fast path -- and hand the kernels the
        # self maps (row i is its own contrast row), which takes the id tests out of the sweeps.  (Small or batched
        # problems are launch-bound: the sort would cost more than it saves.)
        n = idx_a.numel()
        order = torch.argsort(lab_b.long(), stable=True)
        same_labels = lab_b is lab_a
        idx_a = idx_b = idx_a[order]
        lab_b = lab_b[order]
        lab_a = lab_b if same_labels else lab_a[order]
        weight = weight[order]
        id_a = id_b = torch.arange(n, device=feat.device, dtype=torch.int32)
        selfcol = selfrow = id_a
    rowmaps = None
    if metas is not None:
        meta_a, meta_b = metas
    else:
        meta_a = ops.pad_meta(lab_a, id_a)
        meta_b = meta_a if (same_rows and lab_b is lab_a) else ops.pad_meta(lab_b, id_b)
    if n_class > 0 and selfcol is None:
        if same_rows:
            selfcol = torch.arange(idx_a.numel(), device=feat.device, dtype=torch.int32)
            selfrow = selfcol
        elif id_bound is not None:
            selfcol, selfrow, tables = _ops.self_maps_bounded(id_a, id_b, int(id_bound))
            n_pix = feat.shape[0] * feat.shape[2] * feat.shape[3]
            if id_a is idx_a and id_b is idx_b and int(id_bound) == n_pix and (idx_a.numel() + idx_b.numel()) * 4 >= n_pix:
                rowmaps = tables          # the ids ARE the pixel indices: the tables are pixel -> row maps
        else:
            selfcol, selfrow = ops.self_maps(id_a, id_b)
    return _P2PLoss.apply(feat, idx_a.contiguous(), idx_b.contiguous(), meta_a, meta_b, weight.float().contiguous(),
                          float(temperature), bool(normalize), bool(same_rows), int(n_class), selfcol, selfrow, int(n_batch),
                          rowmaps)


_INDEX_CACHE = {}          # row-index tensors per geometry: pure functions of the shapes, rebuilt ~20 small torch ops a call


def _cached(key, build):
    hit = _INDEX_CACHE.get(key)
    if hit is None:
        if len(_INDEX_CACHE) > 64:
            _INDEX_CACHE.clear()
        hit = _INDEX_CACHE[key] = build()
    return hit


def _view_major_rows(b: int, v: int, h: int, w: int, ys: torch.Tensor, xs: torch.Tensor, device) -> torch.Tensor:
    """Pixel indices (into the [b*v, c, h, w] memory-order map) of the reference's row order
    ``cat(unbind(features, dim=1), dim=0)`` (utils/loss.py:337-343) restricted to rows ys, cols xs."""
    grid = (ys.view(-1, 1) * w + xs.view(1, -1)).reshape(-1)                      # [hs*ws]
    img = (torch.arange(b, device=device).view(1, -1) * v + torch.arange(v, device=device).view(-1, 1)).reshape(-1)
    return (img.view(-1, 1) * (h * w) + grid.view(1, -1)).reshape(-1)             # (v, b, y, x) order


def _supcon(features, labels, temperature, ys, xs, zero_if_no_foreground=False, n_class=0, geom=None):
    if features.ndim <= 3:                                                          # :334-336
        raise ValueError('`features` needs to be [bsz, n_views, ...],'
                         'at least 4 dimensions are required')
    if features.ndim != 5:
        raise ValueError("slcl SupConLoss expects [bsz, n_views, c, h, w] features")
    b, v, c, h, w = features.shape
    dev = features.device
    fmap = features.reshape(b * v, c, h, w)
    if geom is not None:          # ys / xs are a pure function of `geom`: reuse the index tensors of the previous call
        idx, ids = _cached(("rows", b, v, h, w, geom, str(dev)),
                           lambda: (lambda i: (i, torch.arange(i.numel(), device=dev, dtype=torch.int32)))(
                               _view_major_rows(b, v, h, w, ys(), xs(), dev)))
    else:
        idx = _view_major_rows(b, v, h, w, ys, xs, dev)
        ids = torch.arange(idx.numel(), device=dev, dtype=torch.int32)
    m = idx.numel()
    if labels is not None:
        # {label, pixel index} rows and the foreground weights fg / sum fg (:352-358, :382-384) in three launches;
        # zero_if_no_foreground: LocalConLoss / BlockConLoss early-out (:405-407, :439-440), without a host sync
        meta = _ops.rows_meta(labels.reshape(-1).long(), idx)
        weight, _ = _ops.tile_weights(meta, m, 1, bool(zero_if_no_foreground))
        if n_class > 0:
            return p2p_loss(fmap, idx, idx, None, None, ids, ids, weight, temperature, normalize=False, same_rows=True,
                            n_class=n_class, metas=(meta, meta))
        lab = meta[:m, 0]                 # general sweeps: p2p_loss may sort the rows by label and rebuilds the metadata
        return p2p_loss(fmap, idx, idx, lab, lab, ids, ids, weight, temperature, normalize=False, same_rows=True, n_class=0)
    lab = (ids % (m // v)).to(torch.int32)                                          # :360-361 same pixel, other views
    weight = torch.full((m,), 1.0 / m, device=dev)                                  # :386
    return p2p_loss(fmap, idx, idx, lab, lab, ids, ids, weight, temperature, normalize=False, same_rows=True, n_class=0)


class SupConLoss(nn.Module):
    """Reference utils/loss.py:315-387.  ``contrast_mode`` / ``base_temperature`` are stored and
    unused there too."""

    def __init__(self, temperature=0.07, contrast_mode='all', base_temperature=0.07, n_class=None):
        super().__init__()
        self.temperature = temperature
        self.contrast_mode = contrast_mode
        self.base_temperature = base_temperature
        # extension: 1..8 = labels are class indices in [0, n_class) -> analytic tensor-core sweeps; 0 = general sweeps
        # (any integer labels); None (default) = decided from the labels themselves (_AutoClasses)
        self.n_class = n_class
        self._classes = _AutoClasses(n_class)

    def forward(self, features, labels=None):
        if features.ndim <= 3:
            raise ValueError('`features` needs to be [bsz, n_views, ...],'
                             'at least 4 dimensions are required')
        h, w = features.shape[-2:]
        dev = features.device
        n_class, in_range = self._classes.resolve(labels)
        return _guard(_supcon(features, labels, self.temperature, lambda: torch.arange(h, device=dev),
                              lambda: torch.arange(w, device=dev), n_class=n_class, geom=("all",)), in_range)


class LocalConLoss(nn.Module):
    """Reference utils/loss.py:390-413: stride-subsampled SupCon."""

    def __init__(self, temperature=0.7, stride=4, n_class=None):
        super().__init__()
        self.temp = temperature
        self.supconloss = SupConLoss(temperature=self.temp, n_class=n_class)
        self.stride = stride

    def forward(self, features, labels=None):
        h, w = features.shape[-2:]
        dev = features.device
        st = self.stride
        n_class, in_range = self.supconloss._classes.resolve(labels)
        return _guard(_supcon(features, labels, self.temp, lambda: torch.arange(0, h, st, device=dev),
                              lambda: torch.arange(0, w, st, device=dev), zero_if_no_foreground=True, n_class=n_class,
                              geom=("stride", st)), in_range)


class BlockConLoss(nn.Module):
    """Reference utils/loss.py:416-466: mean of SupCon over block_size x block_size tiles; tiles whose
    labels are all background are skipped (decided on the device, no per-tile host sync)."""

    def __init__(self, temperature=0.7, block_size=32, n_class=None):
        super().__init__()
        self.block_size = block_size
        self.supconloss = SupConLoss(temperature=temperature, n_class=n_class)
        self.batched = True             # all tiles as one block-diagonal problem where the shape allows (False: per-tile loop)

    def forward(self, features, labels=None):
        dev = features.device
        bs = self.block_size
        div = features.shape[-1] // bs                                               # :426-428
        t = self.supconloss.temperature
        if div == 0:
            return torch.zeros((), device=dev)
        n_class, in_range = self.supconloss._classes.resolve(labels)
        if self.batched and features.ndim == 5:
            b, v = features.shape[:2]
            if (b * v * bs * bs) % 128 == 0:
                return _guard(self._forward_batched(features, labels, div, t, n_class), in_range)
        losses, flags = [], []
        for i in range(div):
            for j in range(div):
                ys = torch.arange(i * bs, (i + 1) * bs, device=dev)
                xs = torch.arange(j * bs, (j + 1) * bs, device=dev)
                losses.append(_supcon(features, labels, t, ys, xs, zero_if_no_foreground=True, n_class=n_class))
                if labels is not None:
                    flags.append((labels[:, :, i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] != 0).any())
        stacked = torch.stack(losses)
        if labels is None:
            return stacked.mean()                                                    # :465 (general sweeps: no guard)
        keep = torch.stack(flags).float()
        return _guard((stacked * keep).sum() / keep.sum().clamp_min(1.0), in_range)   # :445-448 (0 when every tile is skipped)

    def _forward_batched(self, features, labels, div, t, n_class=0):
        """All div x div tiles as ONE block-diagonal problem (one launch per tensor-core sweep instead of div^2
        separate SupCon calls): rows are ordered (tile, view, image, y, x); the per-tile weighting of :445-448 /
        :465 -- fg_i / sum_tile fg, averaged over the tiles that have foreground -- is folded into the row weights.
        ``n_class`` > 0 (class-index labels): the analytic sweeps with one table of per-class sums per tile."""
        b, v, c, h, w = features.shape
        dev = features.device
        bs = self.block_size
        fmap = features.reshape(b * v, c, h, w)

        def build():
            ar = torch.arange(bs, device=dev)
            ti = torch.arange(div, device=dev)
            ys = (ti.view(-1, 1) * bs + ar.view(1, -1))                                  # [div, bs]
            grid = (ys.view(div, 1, bs, 1) * w + ys.view(1, div, 1, bs)).reshape(div * div, bs * bs)      # [tiles, bs*bs]
            img = (torch.arange(b, device=dev).view(1, -1) * v + torch.arange(v, device=dev).view(-1, 1)).reshape(-1)   # (v, b)
            i = (img.view(1, -1, 1) * (h * w) + grid.view(div * div, 1, bs * bs)).reshape(-1)           # (tile, v, b, y, x)
            return i, torch.arange(i.numel(), device=dev, dtype=torch.int32)
        idx, ids = _cached(("blocks", b, v, h, w, bs, div, str(dev)), build)
        n_tiles, m_tile = div * div, b * v * bs * bs
        if labels is not None:
            meta = _ops.rows_meta(labels.reshape(-1).long(), idx)
            weight, _ = _ops.tile_weights(meta, idx.numel(), n_tiles, True)
            return p2p_loss(fmap, idx, idx, None, None, ids, ids, weight, t, normalize=False, same_rows=True,
                            n_class=n_class, n_batch=n_tiles, metas=(meta, meta))
        lab = (ids % (m_tile // v)).to(torch.int32)                                   # same pixel of the tile, other views
        weight = torch.full((idx.numel(),), 1.0 / (m_tile * n_tiles), device=dev)
        return p2p_loss(fmap, idx, idx, lab, lab, ids, ids, weight, t, normalize=False, same_rows=True, n_batch=n_tiles)


def sample_class_balanced(labels: torch.Tensor, n_anchor: int, n_contrast: int, n_class: int,
                          generator: Optional[torch.Generator] = None, return_counts: bool = False):
    """Class-balanced sampler (north_star item 1), with NO host synchronisation.

    ONE permutation of all pixels is drawn from PyTorch's RNG stream -- ``torch.randperm(N, generator)``, on the label
    map's device unless the generator lives elsewhere -- and the labels are read in that order.  The stable per-class
    compaction of the permuted labels (CUDA kernel ``slcl_compact_by_class``, order bit-identical to ``torch.nonzero``)
    gives every pixel its rank within its class; class k contributes its first ``ceil(n / K)`` pixels in permutation
    order (class-major output), and slots a short class leaves open are filled with the next unpicked labelled pixels
    (the test-side restatement of this rule is ``sample_class_balanced`` of the CPU checker).  Output sizes are ``K * ceil(n / K)`` whatever the class
    counts, so nothing has to come back to the host.  Anchors use the same rule with their own quota (their per-class
    picks are a prefix of the contrast picks of that class).  Returns (anchor_idx, contrast_idx[, n_filled_anchor, n_filled_contrast])."""
    lab = labels.reshape(-1).long()
    n = lab.numel()
    dev = lab.device
    gen_dev = generator.device if generator is not None else dev
    perm = torch.randperm(n, generator=generator, device=gen_dev).to(dev)
    per_a = -(-n_anchor // n_class)
    per_c = -(-n_contrast // n_class)
    # labels in permutation order -> stable per-class compaction -> phase-1 prefix per class -> phase-2 fill, both quotas:
    # one C-ABI call (slcl_sample_balanced), ~13 launches where the torch formulation of the same rule took ~60
    a_idx, a_fill, c_idx, c_fill = _ops.sample_balanced(perm, lab, n_class, per_a, per_c)
    a_fill, c_fill = a_fill[0], c_fill[0]
    if return_counts:
        return a_idx, c_idx, a_fill, c_fill
    return a_idx, c_idx


def sampled_supcon_loss(feat: torch.Tensor, labels: torch.Tensor, n_anchor: int, n_contrast: int, n_class: int,
                        temperature: float = 0.7, generator: Optional[torch.Generator] = None,
                        anchor_idx: Optional[torch.Tensor] = None, contrast_idx: Optional[torch.Tensor] = None,
                        analytic: bool = True):
    """Sampled rectangular pixel <-> pixel loss (BASELINE.json configs[2]): anchors x contrast rows
    drawn class-balanced from an NCHW map, L2-normalised, bf16 tensor-core similarities.  The whole call -- draw,
    compaction, gather, sweeps -- runs without a host synchronisation; if the map holds fewer labelled pixels than the
    sample needs, the loss is NaN (decided on the device)."""
    short = None
    if anchor_idx is None or contrast_idx is None:
        anchor_idx, contrast_idx, a_fill, c_fill = sample_class_balanced(labels, n_anchor, n_contrast, n_class, generator,
                                                                          return_counts=True)
        short = (a_fill < anchor_idx.numel()) | (c_fill < contrast_idx.numel())
    lab = labels.reshape(-1).long()
    anchor_idx, contrast_idx = anchor_idx.reshape(-1).long(), contrast_idx.reshape(-1).long()
    meta_a, meta_b = _ops.rows_meta(lab, anchor_idx), _ops.rows_meta(lab, contrast_idx)          # {label, pixel index}
    weight, _ = _ops.tile_weights(meta_a, anchor_idx.numel(), 1, False)          # fg / sum fg over the anchors
    # analytic sweeps need class-index labels (true here) and unique picks (distinct pixels of one permutation are)
    loss = p2p_loss(feat, anchor_idx, contrast_idx, None, None, anchor_idx, contrast_idx, weight, temperature, normalize=True,
                    n_class=n_class if (analytic and n_class <= 8) else 0, metas=(meta_a, meta_b), id_bound=lab.numel())
    if short is not None:
        loss = torch.where(short, torch.full_like(loss, float("nan")), loss)
    return loss


class InterpolatedSupervisedContrastiveLoss(nn.Module):
    """ISCL on [N, d] feature vectors (reference utils/losses.py:6-81): mix-up supervised
    contrastive loss -- two label sets over the SAME Gram matrix, mixed with lambda.  Each term is one sweep of
    the tensor-core kernel with different row (query) labels and the dominant labels on the columns."""

    def __init__(self, temperature):
        super().__init__()
        self.temperature = temperature

    def forward(self, features, labels_1, labels_2, dominant_labels, lambdas, normalize=True):
        n, d = features.shape
        dev = features.device
        fmap = features.reshape(n, d, 1, 1)                 # rows [N,d] are an NCHW map with HW = 1
        idx = torch.arange(n, device=dev)
        ids = idx.to(torch.int32)
        lam = lambdas.to(torch.float32).reshape(-1)
        dom = dominant_labels.reshape(-1)
        t1 = p2p_loss(fmap, idx, idx, labels_1.reshape(-1), dom, ids, ids, lam / n, self.temperature, normalize, same_rows=True)
        t2 = p2p_loss(fmap, idx, idx, labels_2.reshape(-1), dom, ids, ids, (1.0 - lam) / n, self.temperature, normalize,
                      same_rows=True)
        return t1 + t2                                       # :55-59 (losses.mean())
