"""Pixel <-> pixel supervised contrastive losses (reference utils/loss.py:315-466,
duplicated at utils/losses.py:95-239).  Placeholder module body is filled by the
tensor-core path; see p2p kernels."""
from __future__ import annotations

import torch
import torch.nn as nn


class SupConLoss(nn.Module):
    def __init__(self, temperature=0.07, contrast_mode='all', base_temperature=0.07):
        super().__init__()
        self.temperature = temperature
        self.contrast_mode = contrast_mode
        self.base_temperature = base_temperature

    def forward(self, features, labels=None):
        raise NotImplementedError("pixel<->pixel tensor-core path not built yet")


class LocalConLoss(nn.Module):
    def __init__(self, temperature=0.7, stride=4):
        super().__init__()
        self.supconloss = SupConLoss(temperature=temperature)
        self.stride = stride

    def forward(self, features, labels=None):
        raise NotImplementedError("pixel<->pixel tensor-core path not built yet")


class BlockConLoss(nn.Module):
    def __init__(self, temperature=0.7, block_size=32):
        super().__init__()
        self.block_size = block_size
        self.supconloss = SupConLoss(temperature=temperature)

    def forward(self, features, labels=None):
        raise NotImplementedError("pixel<->pixel tensor-core path not built yet")
