"""Actuator parameter containers (``artist/field/actuators*.py``).

The hot path never calls these methods: the motor-position -> joint-angle maps live inside the
kinematics kernels (``csrc/kinematics.cu``).  The methods below restate the same maps with a few
tensor ops for callers that use the actuators on their own (N x 2 values).
"""
from __future__ import annotations

import torch


class Actuators(torch.nn.Module):
    def __init__(self, non_optimizable_parameters: torch.Tensor, optimizable_parameters: torch.Tensor | None = None,
                 device: torch.device | None = None) -> None:
        super().__init__()
        self.non_optimizable_parameters = non_optimizable_parameters
        self.optimizable_parameters = optimizable_parameters if optimizable_parameters is not None else torch.tensor([])
        self.active_non_optimizable_parameters = self.non_optimizable_parameters
        self.active_optimizable_parameters = self.optimizable_parameters

    is_linear = False

    def motor_positions_to_angles(self, motor_positions, device=None):
        raise NotImplementedError("Must be overridden!")

    def angles_to_motor_positions(self, angles, device=None):
        raise NotImplementedError("Must be overridden!")

    def forward(self, motor_positions, device=None):
        return self.motor_positions_to_angles(motor_positions, device)


class IdealActuators(Actuators):
    """Angles == motor positions (``actuators_ideal.py:66-111``)."""

    def motor_positions_to_angles(self, motor_positions, device=None):
        return motor_positions

    def angles_to_motor_positions(self, angles, device=None):
        return angles


class LinearActuators(Actuators):
    """Law-of-cosines linear actuators (``actuators_linear.py:79-370``)."""

    is_linear = True
    epsilon = 1e-6

    def _params(self):
        sp = lambda x: torch.nn.functional.softplus(x, beta=100) + self.epsilon
        no, op = self.active_non_optimizable_parameters, self.active_optimizable_parameters
        return sp(no[:, 4]), sp(no[:, 5]), sp(no[:, 6]), op[:, 0], sp(op[:, 1]), no[:, 1]

    def _absolute_angles(self, motor_positions):
        inc, off, rad, _, s0, _ = self._params()
        stroke = torch.clamp(motor_positions / inc + s0, min=(off - rad).abs() + self.epsilon,
                             max=off + rad - self.epsilon)
        div = (off**2 + rad**2 - stroke**2) / (2.0 * off * rad)
        return torch.arccos(torch.clamp(div, min=-1.0 + 1e-6, max=1.0 - 1e-6))

    def motor_positions_to_angles(self, motor_positions, device=None):
        _, _, _, a0, _, cw = self._params()
        delta = self._absolute_angles(torch.zeros_like(motor_positions)) - self._absolute_angles(motor_positions)
        return a0 + delta * (cw == 1) - delta * (cw == 0)

    def angles_to_motor_positions(self, angles, device=None):
        inc, off, rad, a0, s0, cw = self._params()
        delta = torch.where(cw == 1, angles - a0, a0 - angles)
        ia = self._absolute_angles(torch.zeros_like(angles)) - delta
        cosv = torch.clamp(torch.cos(ia), -1.0 + 1e-6, 1.0 - 1e-6)
        stroke = torch.sqrt(off**2 + rad**2 - 2.0 * off * rad * cosv)
        stroke = torch.clamp(stroke, min=(off - rad).abs() + self.epsilon, max=off + rad - self.epsilon)
        return (stroke - s0) * inc


actuator_type_mapping = {0: LinearActuators, 1: IdealActuators}
