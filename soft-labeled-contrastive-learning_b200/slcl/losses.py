"""Drop-in for the reference's ``utils/losses.py``: the SupCon family (:95-239 duplicate ``utils/loss.py``
verbatim) and the mix-up ISCL loss (:6-81)."""
from .p2p import SupConLoss, LocalConLoss, BlockConLoss, InterpolatedSupervisedContrastiveLoss  # noqa: F401
