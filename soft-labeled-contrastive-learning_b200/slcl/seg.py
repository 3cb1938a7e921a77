"""Segmentation losses next to the contrastive path: drop-ins for the reference's
``loss_calc`` / ``jaccard_loss`` / ``dice_loss`` (utils/loss.py:11-103) and ``prob_2_entropy``
(utils/utils_.py:627-631).  One fused CUDA pass over the logits computes cross entropy, Dice and Jaccard
together (``seg_losses``); the drop-in functions pick what they need from it."""
from __future__ import annotations

import torch

from . import ops  # noqa: F401

_ops = ops.dispatch          # eager: op bodies directly; compiled: torch.ops.slcl


class _SegLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        losses, stats = _ops.seg_fwd(logits.detach(), labels)
        ctx.save_for_backward(logits, labels, stats)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        logits, labels, stats = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None
        return _ops.seg_bwd(logits.detach(), labels, stats, grad_losses.contiguous()), None


def seg_losses(pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """[3] tensor {cross entropy, dice loss, jaccard loss} of logits ``pred`` [B,K,H,W] vs ``label`` [B,H,W]
    (or [B,1,H,W]) in one pass; differentiable w.r.t. ``pred``."""
    if pred.shape[1] < 2:
        raise NotImplementedError("num_classes == 1 (sigmoid branch of jaccard_loss, utils/loss.py:23-32) is not built")
    return _SegLosses.apply(pred, label.to(pred.device))


def loss_calc(pred, label, gpu=0, jaccard=False):
    """utils/loss.py:46-66: CrossEntropyLoss (+ jaccard_loss)."""
    out = seg_losses(pred, label)
    return out[0] + out[2] if jaccard else out[0]


def jaccard_loss(true, logits, eps=1e-7):
    """utils/loss.py:11-43 (eps is fixed at the reference default 1e-7)."""
    if eps != 1e-7:
        raise NotImplementedError("jaccard_loss: only the reference default eps=1e-7 is built")
    return seg_losses(logits, true)[2]


def dice_loss(pred, target):
    """utils/loss.py:69-103."""
    return seg_losses(pred, target)[1]


class _EntropyMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob):
        ctx.save_for_backward(prob)
        return _ops.entropy_map(prob.detach())

    @staticmethod
    def backward(ctx, grad_out):
        (prob,) = ctx.saved_tensors
        return _ops.entropy_map_bwd(prob.detach(), grad_out.contiguous())


def prob_2_entropy(prob):
    """utils/utils_.py:627-631: weighted self-information map -p log2(p + 1e-7) / log2(C) of [N,C,H,W] probabilities."""
    return _EntropyMap.apply(prob)
