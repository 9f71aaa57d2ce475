from .geometry import line_cylinder_intersections, line_plane_intersections, reflect
from .heliostat_ray_tracer import HeliostatRayTracer
from .sampling import DistortionsDataset, RestrictedDistributedSampler

__all__ = ["HeliostatRayTracer", "DistortionsDataset", "RestrictedDistributedSampler", "reflect",
           "line_cylinder_intersections", "line_plane_intersections"]
