from .heliostat_ray_tracer import HeliostatRayTracer
from .sampling import DistortionsDataset, RestrictedDistributedSampler

__all__ = ["HeliostatRayTracer", "DistortionsDataset", "RestrictedDistributedSampler"]
