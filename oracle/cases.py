"""Seeded input recipes shared by oracle/make_golden.py and tests/ (TEST
INFRASTRUCTURE).  Inputs are always generated on the CPU from
``torch.Generator().manual_seed(seed)`` so the oracle, the committed golden
outputs and the CUDA path see identical bits.
The KAT recipes are the ones of SURVEY.md section 8(c).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def g(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed)


def shipped_centres() -> torch.Tensor:
    """The (4,32) fp32 class-centre state.  The .npy shipped by the reference
    (class_center_ct_f0.npy) is 640 bytes of *data*; a byte-identical copy of
    its payload is kept in tests/golden/class_center_ct_f0.npy as a fixture."""
    return torch.from_numpy(np.load(os.path.join(GOLDEN_DIR, "class_center_ct_f0.npy"))).float()


def kat1():
    feas = torch.randn(2, 32, 8, 8, generator=g(1))
    labels = torch.randint(0, 4, (2, 8, 8), generator=g(2))
    return feas, labels


def kat2():
    return torch.randn(2, 32, 8, 8, generator=g(3))


def kat4():
    return torch.randn(4, 32, generator=g(4)), torch.randn(4, 32, generator=g(5))


def kat6():
    f5 = F.normalize(torch.randn(1, 2, 32, 8, 8, generator=g(6)), dim=2)
    lab = torch.randint(0, 4, (1, 2, 8, 8), generator=g(7))
    return f5, lab


def kat7():
    f7 = F.normalize(torch.randn(1, 2, 32, 64, 64, generator=g(8)), dim=2)
    lab7 = torch.randint(0, 4, (1, 2, 64, 64), generator=g(9))
    return f7, lab7


def kat8_mask():
    return torch.softmax(torch.randn(128, 4, generator=g(10)), 1)


def kat9():
    gen = g(11)
    feas = torch.randn(2, 128, 33, 33, generator=gen)
    lab = torch.randint(0, 5, (2, 256, 256), generator=gen)
    cc = torch.randn(5, 128, generator=gen)
    return feas, lab, cc


def soft_case(seed: int = 21, b: int = 2, c: int = 32, h: int = 12, w: int = 10, k: int = 4):
    """Soft-label centroid inputs: features + softmax(3*N(0,1)) probabilities."""
    gen = g(seed)
    ft = torch.randn(b, c, h, w, generator=gen)
    probs = torch.softmax(3.0 * torch.randn(b, k, h, w, generator=gen), dim=1)
    return ft, probs


def ragged_case(seed: int = 31, b: int = 3, c: int = 20, h: int = 7, w: int = 9, k: int = 5):
    """Odd sizes (HW not a multiple of 4, C not a multiple of 8), labels with
    out-of-range values (-1 and K) and one empty class."""
    gen = g(seed)
    ft = torch.randn(b, c, h, w, generator=gen)
    lab = torch.randint(0, k - 1, (b, h, w), generator=gen)       # class k-1 never occurs
    lab[0, 0, 0] = -1
    lab[1, 2, 3] = k
    cc = torch.randn(k, c, generator=gen)
    sel = (torch.rand(b * h * w, generator=gen) > 0.4).float()
    return ft, lab, cc, sel


def seg_case(seed: int = 51, b: int = 3, k: int = 4, h: int = 12, w: int = 10):
    """Segmentation logits + labels; class k-1 is absent from image 0."""
    gen = g(seed)
    logits = 2.0 * torch.randn(b, k, h, w, generator=gen)
    labels = torch.randint(0, k, (b, h, w), generator=gen)
    labels[0][labels[0] == k - 1] = 0
    return logits, labels


def iscl_case(seed: int = 71, n: int = 96, d: int = 48, k: int = 4):
    """Mix-up batch for ISCL: features, two label sets, dominant labels, lambdas."""
    gen = g(seed)
    feats = torch.randn(n, d, generator=gen)
    l1 = torch.randint(0, k, (n,), generator=gen)
    l2 = l1[torch.randperm(n, generator=gen)]
    lam = torch.rand(n, generator=gen)
    dom = torch.where(lam >= 0.5, l1, l2)
    return feats, l1, l2, dom, lam
