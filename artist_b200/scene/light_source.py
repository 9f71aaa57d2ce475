import torch


class LightSource:
    """Abstract light source (``artist/scene/light_source.py:6-97``)."""

    def __init__(self, number_of_rays: int) -> None:
        self.number_of_rays = number_of_rays

    def get_distortions(self, number_of_points: int, number_of_active_heliostats: int,
                        random_seed: int = 7) -> tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError("Must be overridden!")
