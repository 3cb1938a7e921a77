"""Class-centre state in the reference's on-disk layout:
NPY v1.0, little-endian fp32, C order, shape (K, C) -- e.g. the shipped
``class_center_ct_f0.npy`` is (4, 32), 128-byte header + 512-byte payload.
Loader mirrors trainer/Trainer_MPSCL.py:306-323."""
from __future__ import annotations

import os

import numpy as np
import torch


def class_center_filename(data_dir: str, fold: int) -> str:
    """``class_center_<bssfp|ct>_f<fold>.npy`` chosen like Trainer_MPSCL.py:306-307."""
    modality = "bssfp" if "mscmrseg" in data_dir else "ct"
    return f"class_center_{modality}_f{fold}.npy"


def load_class_centers(path: str, device="cuda") -> torch.Tensor:
    arr = np.load(path)
    if arr.ndim != 2:
        raise ValueError(f"{path}: expected a [K, C] array, got shape {arr.shape}")
    return torch.from_numpy(np.ascontiguousarray(arr, dtype="<f4")).float().to(device)


def save_class_centers(path: str, centres: torch.Tensor) -> None:
    arr = np.ascontiguousarray(centres.detach().float().cpu().numpy(), dtype="<f4")
    if arr.ndim != 2:
        raise ValueError("class centres must be [K, C]")
    tmp = path + ".tmp"
    with open(tmp, "wb") as fh:
        np.lib.format.write_array(fh, arr, version=(1, 0))
    os.replace(tmp, path)


# ---- f-3: the reference checkpoints {'epoch','model_state_dict','optimizer_state_dict'}
# (utils/callbacks.py:67-69) and silently loses the class-centre state on resume (MPSCL reloads the shipped .npy,
# trainer/Trainer_MPSCL.py:306-323; MCCL resets centroid_s every epoch, trainer/Trainer_MCCL.py:179).
CENTRE_KEY = "class_center_feas"


def add_to_checkpoint(checkpoint: dict, centres: torch.Tensor, key: str = CENTRE_KEY) -> dict:
    """Put the [K,C] class-centre state into a reference-style checkpoint dict (fp32, CPU, C order)."""
    if centres.dim() != 2:
        raise ValueError("class centres must be [K, C]")
    checkpoint[key] = centres.detach().float().cpu().contiguous()
    return checkpoint


def from_checkpoint(checkpoint: dict, device="cuda", key: str = CENTRE_KEY, fallback_npy: str = None) -> torch.Tensor:
    """Class centres from a checkpoint written with ``add_to_checkpoint``; checkpoints of the unmodified
    reference have no such key, then ``fallback_npy`` (the shipped class_center_*.npy) is loaded instead."""
    if key in checkpoint:
        return checkpoint[key].float().to(device)
    if fallback_npy is None:
        raise KeyError(f"checkpoint has no '{key}' and no fallback .npy was given")
    return load_class_centers(fallback_npy, device=device)
