from .surfaces import NURBSSurfaces
from .utils import create_nurbs_evaluation_grid, create_planar_nurbs_control_points

__all__ = ["NURBSSurfaces", "create_nurbs_evaluation_grid", "create_planar_nurbs_control_points"]
