from __future__ import annotations

from collections.abc import Iterator

import torch


class DistortionsDataset:
    """Pre-sampled sun-shape distortions for every active heliostat-sample
    (``artist/raytracing/sampling.py:10-85``): ``distortions_u/_e`` are ``[N, R, P]``."""

    def __init__(self, light_source, number_of_points_per_heliostat: int, number_of_active_heliostats: int,
                 random_seed: int = 7) -> None:
        self.distortions_u, self.distortions_e = light_source.get_distortions(
            number_of_points=number_of_points_per_heliostat,
            number_of_active_heliostats=number_of_active_heliostats, random_seed=random_seed)

    def __len__(self) -> int:
        return self.distortions_u.shape[0]

    def __getitem__(self, idx: int) -> tuple[torch.Tensor, torch.Tensor]:
        return self.distortions_u[idx], self.distortions_e[idx]


class RestrictedDistributedSampler:
    """Sharding contract (``artist/raytracing/sampling.py:88-157``): heliostat ``h`` with all of its
    replicated samples (contiguous rows) belongs to rank ``h % min(n_heliostats, world_size)``; ranks
    beyond the number of heliostats stay idle instead of receiving duplicated rays."""

    def __init__(self, number_of_samples: int, number_of_active_heliostats: int, world_size: int = 1,
                 rank: int = 0) -> None:
        number_of_active_heliostats = int(number_of_active_heliostats)
        active_ranks = min(number_of_active_heliostats, world_size)
        self.rank_indices: list[int] = []
        if rank < active_ranks:
            per_heliostat = number_of_samples // number_of_active_heliostats
            for h in range(rank, number_of_active_heliostats, active_ranks):
                self.rank_indices.extend(range(h * per_heliostat, (h + 1) * per_heliostat))

    def __iter__(self) -> Iterator[int]:
        return iter(self.rank_indices)

    def __len__(self) -> int:
        return len(self.rank_indices)
