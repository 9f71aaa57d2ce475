from __future__ import annotations

import torch

from .. import ops
from .heliostat_group import HeliostatGroup
from .kinematics_rigid_body import RigidBody


class HeliostatGroupRigidBody(HeliostatGroup):
    """Heliostat group with rigid-body kinematics (``artist/field/heliostat_group_rigid_body.py:13-270``)."""

    def __init__(self, names, positions, surface_points, surface_normals, canting, facet_translations,
                 initial_orientations, nurbs_control_points, nurbs_degrees, kinematics_translation_deviation_parameters,
                 kinematics_rotation_deviation_parameters, actuator_parameters_non_optimizable,
                 actuator_parameters_optimizable: torch.Tensor = torch.tensor([]), device=None) -> None:
        super().__init__(names=names, positions=positions, surface_points=surface_points,
                         surface_normals=surface_normals, canting=canting, facet_translations=facet_translations,
                         initial_orientations=initial_orientations, nurbs_control_points=nurbs_control_points,
                         nurbs_degrees=nurbs_degrees, device=device)
        self.kinematics = RigidBody(
            number_of_heliostats=self.number_of_heliostats, heliostat_positions=self.positions,
            initial_orientations=self.initial_orientations,
            translation_deviation_parameters=kinematics_translation_deviation_parameters,
            rotation_deviation_parameters=kinematics_rotation_deviation_parameters,
            actuator_parameters_non_optimizable=actuator_parameters_non_optimizable,
            actuator_parameters_optimizable=actuator_parameters_optimizable, device=device)

    def _apply(self, orientations: torch.Tensor) -> None:
        # points @ O^T, normals @ O^T (:217-222, :265-270) - recorded, not executed: the ray tracer fuses the rotation
        # into its kernels; reading active_surface_points / _normals materialises them (HeliostatGroup properties)
        self._record_alignment(orientations)

    def align_surfaces_with_incident_ray_directions(self, aim_points, incident_ray_directions, active_heliostats_mask,
                                                    device=None) -> None:
        assert torch.equal(self.active_heliostats_mask, active_heliostats_mask), (
            "Some heliostats were not activated and cannot be aligned.")
        self._apply(self.kinematics.incident_ray_directions_to_orientations(
            incident_ray_directions=incident_ray_directions, aim_points=aim_points))

    def align_surfaces_with_motor_positions(self, motor_positions, active_heliostats_mask, device=None) -> None:
        assert torch.equal(self.active_heliostats_mask, active_heliostats_mask), (
            "Some heliostats were not activated and cannot be aligned.")
        self._apply(self.kinematics.motor_positions_to_orientations(motor_positions=motor_positions))
