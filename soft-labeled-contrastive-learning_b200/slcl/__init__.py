"""slcl -- B200-native (sm_100a) implementation of the SLCL contrastive-loss hot
path, drop-in for the loss API of Dinhthixuanbinh/Soft-Labeled-Contrastive-Learning.

    from slcl.loss import MPCL, mpcl_loss_calc, ContrastiveLoss, SupConLoss, LocalConLoss, BlockConLoss
    from slcl.utils_ import cal_centroid, update_class_center_iter, generate_pseudo_label

replace ``from utils.loss import ...`` / ``from utils.utils_ import ...`` in the
reference trainers (trainer/Trainer_MPSCL.py:14-15, trainer/Trainer_MCCL.py:14,17).
All arithmetic runs in hand-written CUDA kernels behind the C ABI of
``libslcl.so`` (include/slcl.h); there is no CPU fallback.
"""
from ._lib import SlclError, LIB_PATH, load as load_library  # noqa: F401

__version__ = "1.0"
