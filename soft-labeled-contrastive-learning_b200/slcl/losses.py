"""Drop-in for the pixel <-> pixel part of the reference's ``utils/losses.py``
(:95-239 duplicate the SupCon family of ``utils/loss.py`` verbatim)."""
from .p2p import SupConLoss, LocalConLoss, BlockConLoss  # noqa: F401
