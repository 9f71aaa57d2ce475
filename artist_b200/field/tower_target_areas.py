"""Target-area SoA containers (``artist/field/tower_target_areas*.py``)."""
import torch


class TowerTargetAreas:
    def __init__(self, names: list[str], centers: torch.Tensor, normals: torch.Tensor) -> None:
        self.names = names
        self.centers = centers
        self.normals = normals
        self.number_of_target_areas = len(names)


class TowerTargetAreasPlanar(TowerTargetAreas):
    """Planar areas: ``centers [T,4]``, ``normals [T,4]``, ``dimensions [T,2]`` (width e, height u)."""

    def __init__(self, names, centers, normals, dimensions) -> None:
        super().__init__(names, centers, normals)
        self.dimensions = dimensions


class TowerTargetAreasCylindrical(TowerTargetAreas):
    """Cylindrical sectors: + ``axes [T,4]``, ``radii``, ``heights``, ``opening_angles [T]``."""

    def __init__(self, names, centers, normals, axes, radii, heights, opening_angles) -> None:
        super().__init__(names, centers, normals)
        self.axes = axes
        self.radii = radii
        self.heights = heights
        self.opening_angles = opening_angles
