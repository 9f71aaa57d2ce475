"""Bitmap losses with the reference's class API (``artist/optim/loss.py``): ``PixelLoss`` (:251-319) and
``KLDivergenceLoss`` (:322-410), each ONE fused CUDA kernel forward and one backward (``csrc/flux.cu``) instead of ~10
eager passes over the ``[N,U,E]`` tensors.  Same call signature, keyword contract and error messages; the reduction must
cover the whole bitmap (``reduction_dimensions=(1, 2)``, what every caller of the reference passes)."""
from __future__ import annotations

from typing import Any

import torch

from .. import ops


class Loss:
    """Abstract base (``loss.py:12-58``)."""

    def __init__(self, loss_function: Any = None) -> None:
        self.loss_function = loss_function

    def __call__(self, prediction: torch.Tensor, ground_truth: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        raise NotImplementedError("Must be overridden!")


def _whole_bitmap(reduction_dimensions) -> None:
    dims = tuple(sorted(int(d) % 3 for d in reduction_dimensions))
    if dims != (1, 2):
        raise ValueError(f"artist_b200 bitmap losses reduce over the whole bitmap: reduction_dimensions must be (1, 2), got "
                         f"{tuple(reduction_dimensions)}")


class PixelLoss(Loss):
    """Summed squared pixel error divided by the total ground-truth intensity of the sample (``loss.py:251-319``)."""

    def __init__(self) -> None:
        super().__init__(loss_function=None)

    def __call__(self, prediction: torch.Tensor, ground_truth: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        expected_kwargs = ["reduction_dimensions"]
        errors = [f"Please add '{key}' as keyword argument." for key in expected_kwargs if key not in kwargs]
        if errors:
            raise ValueError(f"The vector loss expects {expected_kwargs} as keyword arguments. " + " ".join(errors))
        _whole_bitmap(kwargs["reduction_dimensions"])
        return ops.flux_loss(prediction, ground_truth, ops.LOSS_PIXEL)


class KLDivergenceLoss(Loss):
    """``D_KL(P || Q)`` of the L1-normalised ground truth ``P`` and prediction ``Q`` (``loss.py:322-410``)."""

    def __init__(self) -> None:
        super().__init__(loss_function=None)

    def __call__(self, prediction: torch.Tensor, ground_truth: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        for key in ["reduction_dimensions"]:
            if key not in kwargs:
                raise ValueError(f"The KL-divergence loss expects '{key}' as keyword argument. Please add this argument.")
        _whole_bitmap(kwargs["reduction_dimensions"])
        return ops.flux_loss(prediction, ground_truth, ops.LOSS_KL_DIVERGENCE)
