from __future__ import annotations

import torch

from .. import ops
from ..util.env import get_device


class NURBSSurfaces(torch.nn.Module):
    """Batched NURBS facets (``artist/nurbs/surfaces.py:8-727``) evaluated by ``ab200_nurbs_fwd`` /
    differentiated by ``ab200_nurbs_bwd``.

    ``control_points`` is ``[N, F, cu, cv, 3]``; knot vectors are clamped-uniform (``:98-155``).
    ``uniform=False`` only changes how the reference searches spans over the same uniform knots,
    so both settings share one kernel.
    """

    def __init__(self, degrees: torch.Tensor, control_points: torch.Tensor, uniform: bool = True,
                 device: torch.device | None = None) -> None:
        super().__init__()
        device = get_device(device)
        self.degrees = degrees
        self.control_points = control_points
        self.uniform = uniform
        self.number_of_surfaces = control_points.shape[0]
        self.number_of_facets_per_surface = control_points.shape[1]
        self._degree_u, self._degree_v = int(degrees[0]), int(degrees[1])
        self._knots_u = self._uniform_knots(control_points.shape[2], self._degree_u, device)
        self._knots_v = self._uniform_knots(control_points.shape[3], self._degree_v, device)

    @staticmethod
    def _uniform_knots(n_ctrl: int, degree: int, device) -> torch.Tensor:
        knots = torch.zeros(n_ctrl + degree + 1, device=device)
        knots[degree:-degree] = torch.linspace(0, 1, n_ctrl - degree + 1, device=device)
        knots[-degree:] = 1
        return knots

    def calculate_uniform_knot_vectors(self, direction: int, device: torch.device | None = None) -> torch.Tensor:
        k = self._knots_u if direction == 0 else self._knots_v
        return k.unsqueeze(0).repeat(self.number_of_surfaces, self.number_of_facets_per_surface, 1)

    @property
    def knot_vectors_u(self) -> torch.Tensor:
        return self.calculate_uniform_knot_vectors(0)

    @property
    def knot_vectors_v(self) -> torch.Tensor:
        return self.calculate_uniform_knot_vectors(1)

    def calculate_surface_points_and_normals(self, evaluation_points: torch.Tensor, canting: torch.Tensor | None,
                                             facet_translations: torch.Tensor | None,
                                             device: torch.device | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        """``evaluation_points [N,F,K,2]`` -> points, normals ``[N,F,K,4]`` (``:475-689``)."""
        return ops.nurbs_points_and_normals(self.control_points, evaluation_points, self._knots_u, self._knots_v,
                                            self._degree_u, self._degree_v, canting, facet_translations)

    def forward(self, evaluation_points, canting, facet_translations, device=None):
        return self.calculate_surface_points_and_normals(evaluation_points, canting, facet_translations, device)
