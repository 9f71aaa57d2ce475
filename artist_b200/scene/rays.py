import torch


class Rays:
    """Ray bundle container (``artist/scene/rays.py:4-50``): directions ``[...,4]`` + magnitudes ``[...]``."""

    def __init__(self, ray_directions: torch.Tensor, ray_magnitudes: torch.Tensor) -> None:
        if ray_directions.shape[-1] != 4 or ray_directions.shape[:-1] != ray_magnitudes.shape:
            raise ValueError("Ray directions and magnitudes have incompatible sizes!")
        self.ray_directions = ray_directions
        self.ray_magnitudes = ray_magnitudes
