import torch

from ..util.env import get_device


def create_nurbs_evaluation_grid(number_of_evaluation_points: torch.Tensor, epsilon: float = 1e-7,
                                 device: torch.device | None = None) -> torch.Tensor:
    """Cartesian (u,v) grid in ``(eps, 1-eps)^2`` -> ``[pu*pv, 2]`` (``artist/nurbs/utils.py:7-49``)."""
    device = get_device(device)
    pu, pv = int(number_of_evaluation_points[0]), int(number_of_evaluation_points[1])
    return torch.cartesian_prod(torch.linspace(epsilon, 1 - epsilon, pu, device=device),
                                torch.linspace(epsilon, 1 - epsilon, pv, device=device))


def create_planar_nurbs_control_points(number_of_control_points: torch.Tensor, canting: torch.Tensor,
                                       device: torch.device | None = None) -> torch.Tensor:
    """Flat control net per facet, ``[F, cu, cv, 3]`` (``artist/nurbs/utils.py:52-121``)."""
    device = get_device(device)
    canting = canting.to(device)
    cu, cv = int(number_of_control_points[0]), int(number_of_control_points[1])
    cp = torch.zeros(canting.shape[0], cu, cv, 3, device=device, dtype=canting.dtype)
    dims = torch.norm(canting, dim=2)
    ul = torch.linspace(0, 1, cu, device=device, dtype=canting.dtype)
    vl = torch.linspace(0, 1, cv, device=device, dtype=canting.dtype)
    cp[..., 0] = (-dims[:, 0, None] + 2 * dims[:, 0, None] * ul)[:, :, None]
    cp[..., 1] = (-dims[:, 1, None] + 2 * dims[:, 1, None] * vl)[:, None, :]
    return cp
