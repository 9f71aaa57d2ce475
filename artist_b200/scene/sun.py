from __future__ import annotations

from typing import Any

import torch

from ..util.env import get_device
from .light_source import LightSource


class Sun(LightSource):
    """Sun with a bivariate normal sun shape (``artist/scene/sun.py:16-234``).

    ``get_distortions`` keeps the reference contract exactly - it reseeds the GLOBAL torch RNG and draws
    ``MultivariateNormal(mean, cov*I).sample((N, R, P))`` - because parity is defined on identical
    distortion samples; the two returned tensors are views of one interleaved ``[N,R,P,2]`` buffer,
    which is the layout the trace kernels stream (no repacking).
    """

    def __init__(self, number_of_rays: int,
                 distribution_parameters: dict[str, Any] = dict(distribution_type="normal", mean=0.0,
                                                                covariance=4.3681e-06),
                 device: torch.device | None = None) -> None:
        super().__init__(number_of_rays=number_of_rays)
        device = get_device(device)
        self.distribution_parameters = distribution_parameters
        if distribution_parameters["distribution_type"] != "normal":
            raise ValueError("Unknown sunlight distribution type.")
        mean = torch.tensor([distribution_parameters["mean"]] * 2, dtype=torch.float, device=device)
        cov = distribution_parameters["covariance"]
        covariance = torch.tensor([[cov, 0], [0, cov]], dtype=torch.float, device=device)
        self.distribution = torch.distributions.MultivariateNormal(mean, covariance)

    @property
    def scatter_sigma(self) -> float:
        return float(self.distribution_parameters["covariance"]) ** 0.5

    def get_distortions(self, number_of_points: int, number_of_active_heliostats: int,
                        random_seed: int = 7) -> tuple[torch.Tensor, torch.Tensor]:
        torch.manual_seed(random_seed)
        torch.cuda.manual_seed(random_seed)
        sample = self.distribution.sample((number_of_active_heliostats, self.number_of_rays, number_of_points))
        distortions_u, distortions_e = sample.permute(3, 0, 1, 2)
        return distortions_u, distortions_e
