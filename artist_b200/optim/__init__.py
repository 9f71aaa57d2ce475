"""Losses of the reference's optimisers that run on the ray tracer's flux bitmaps (``artist/optim/loss.py``)."""
from .loss import KLDivergenceLoss, Loss, PixelLoss

__all__ = ["Loss", "PixelLoss", "KLDivergenceLoss"]
