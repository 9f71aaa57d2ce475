"""artist_b200 - B200-native (sm_100a) implementation of ARTIST's differentiable heliostat
ray-tracing hot path behind the reference's own class API.  See DESIGN.md / INTEGRATION.md."""
from .field import (HeliostatField, HeliostatGroup, HeliostatGroupRigidBody, RigidBody, SolarTower,
                    TowerTargetAreasCylindrical, TowerTargetAreasPlanar)
from .nurbs import NURBSSurfaces, create_nurbs_evaluation_grid, create_planar_nurbs_control_points
from .raytracing import HeliostatRayTracer
from .scenario import Scenario, build_synthetic_scenario, synthetic_field_tensors
from .scene import LightSourceArray, Rays, Sun

__version__ = "0.1.0"
__all__ = ["HeliostatField", "HeliostatGroup", "HeliostatGroupRigidBody", "RigidBody", "SolarTower",
           "TowerTargetAreasCylindrical", "TowerTargetAreasPlanar", "NURBSSurfaces", "create_nurbs_evaluation_grid",
           "create_planar_nurbs_control_points", "HeliostatRayTracer", "Scenario", "build_synthetic_scenario",
           "synthetic_field_tensors", "LightSourceArray", "Rays", "Sun"]
