from __future__ import annotations

import torch

from .. import ops
from ..util.env import get_device
from .actuators import Actuators, actuator_type_mapping


class Kinematics(torch.nn.Module):
    """Abstract kinematics (``artist/field/kinematics.py:8-119``)."""

    def incident_ray_directions_to_orientations(self, incident_ray_directions, aim_points, device=None):
        raise NotImplementedError("Must be overridden!")

    def motor_positions_to_orientations(self, motor_positions, device=None):
        raise NotImplementedError("Must be overridden!")

    def forward(self, incident_ray_directions, aim_points, device=None):
        return self.incident_ray_directions_to_orientations(incident_ray_directions, aim_points, device)


def _initial_orientation_offset() -> torch.Tensor:
    """Rotation taking the sampled surface orientation (up) to the kinematics' standard orientation
    (south): ``rotate_e(a_e) @ rotate_n(a_n) @ rotate_u(a_u)`` with the axis-angle components of
    ``artist/geometry/rotations.py:43-64`` (``kinematics_rigid_body.py:178-190``).  Constant, built on the CPU."""
    up = torch.nn.functional.normalize(torch.tensor([[0.0, 0.0, 1.0]]))
    south = torch.nn.functional.normalize(torch.tensor([0.0, -1.0, 0.0]), dim=0).unsqueeze(0)
    axis = torch.nn.functional.normalize(torch.linalg.cross(up, south))
    comp = torch.arccos(torch.clamp(up @ south.T, -1.0, 1.0)) * axis

    def rot(i, j, ang):
        m = torch.eye(4)
        c, s = torch.cos(ang), torch.sin(ang)
        m[i, i], m[i, j], m[j, i], m[j, j] = c, -s, s, c
        return m

    return rot(1, 2, comp[0, 0]) @ rot(0, 2, comp[0, 1]) @ rot(0, 1, comp[0, 2])


class RigidBody(Kinematics):
    """Two-joint rigid-body kinematics (``artist/field/kinematics_rigid_body.py:15-634``); the forward chain,
    its backward and the fixed-point alignment run as the ``ab200_kinematics_*`` kernels."""

    def __init__(self, number_of_heliostats: int, heliostat_positions: torch.Tensor, initial_orientations: torch.Tensor,
                 translation_deviation_parameters: torch.Tensor, rotation_deviation_parameters: torch.Tensor,
                 actuator_parameters_non_optimizable: torch.Tensor,
                 actuator_parameters_optimizable: torch.Tensor = torch.tensor([]),
                 device: torch.device | None = None) -> None:
        super().__init__()
        device = get_device(device)
        self.number_of_heliostats = number_of_heliostats
        self.heliostat_positions = heliostat_positions
        self.initial_orientations = initial_orientations
        self.motor_positions = torch.zeros(number_of_heliostats, 2, device=device)
        self.translation_deviation_parameters = translation_deviation_parameters
        self.rotation_deviation_parameters = rotation_deviation_parameters
        self.number_of_active_heliostats = 0
        self.active_heliostat_positions = heliostat_positions
        self.active_initial_orientations = initial_orientations
        self.active_translation_deviation_parameters = translation_deviation_parameters
        self.active_rotation_deviation_parameters = rotation_deviation_parameters
        self.active_motor_positions = self.motor_positions
        actuator_type = int(actuator_parameters_non_optimizable[0, 0, 0].item())
        self.actuators: Actuators = actuator_type_mapping[actuator_type](
            non_optimizable_parameters=actuator_parameters_non_optimizable,
            optimizable_parameters=actuator_parameters_optimizable.to(device), device=device)
        self.kinematics_standard_orientation = torch.tensor([0.0, -1.0, 0.0, 0.0], device=device)
        self.initial_orientation_offsets = _initial_orientation_offset().unsqueeze(0).to(device)
        self.homogeneous_origin = torch.tensor([0.0, 0.0, 0.0, 1.0], device=device)

    def _offset(self) -> torch.Tensor:
        return self.initial_orientation_offsets.reshape(4, 4)

    def motor_positions_to_orientations(self, motor_positions: torch.Tensor, device=None) -> torch.Tensor:
        """``[N,2]`` motor positions -> ``[N,4,4]`` orientations, differentiable w.r.t. motor positions,
        rotation / translation deviations and the optimizable actuator parameters (``:510-538``)."""
        act = self.actuators
        return ops.kinematics_orientations(
            motor_positions, self.active_rotation_deviation_parameters, self.active_translation_deviation_parameters,
            act.active_optimizable_parameters if act.is_linear else None, self.active_heliostat_positions,
            act.active_non_optimizable_parameters, self._offset(), act.is_linear)

    def incident_ray_directions_to_orientations(self, incident_ray_directions: torch.Tensor, aim_points: torch.Tensor,
                                                device=None, max_num_iterations: int = 4,
                                                min_eps: float = 0.0001) -> torch.Tensor:
        """<= ``max_num_iterations`` forward/inverse kinematics sweeps, stopping when every heliostat
        converged (``:540-634``); stores ``active_motor_positions``.  Forward only, like every caller uses it."""
        act = self.actuators
        orientations, motor = ops.kinematics_align_incident(
            incident_ray_directions, aim_points, self.active_rotation_deviation_parameters,
            self.active_translation_deviation_parameters, act.active_optimizable_parameters if act.is_linear else None,
            self.active_heliostat_positions, act.active_non_optimizable_parameters, self._offset(), act.is_linear,
            max_num_iterations, min_eps)
        self.active_motor_positions = motor
        return orientations
