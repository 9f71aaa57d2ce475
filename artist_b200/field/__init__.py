from .actuators import Actuators, IdealActuators, LinearActuators
from .heliostat_field import HeliostatField
from .heliostat_group import HeliostatGroup
from .heliostat_group_rigid_body import HeliostatGroupRigidBody
from .kinematics_rigid_body import Kinematics, RigidBody
from .solar_tower import SolarTower
from .tower_target_areas import TowerTargetAreas, TowerTargetAreasCylindrical, TowerTargetAreasPlanar

__all__ = ["Actuators", "IdealActuators", "LinearActuators", "HeliostatField", "HeliostatGroup",
           "HeliostatGroupRigidBody", "Kinematics", "RigidBody", "SolarTower", "TowerTargetAreas",
           "TowerTargetAreasCylindrical", "TowerTargetAreasPlanar"]
