"""Bit-identity of the kinematics kernels against the oracle (forward from identical motor positions, and the
alignment loop).  GPU box: python tools/diag_kinematics_parity.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import artist_oracle as O
from tests import cases
from tests.test_gpu_ops_parity import _kin_dev, DEV
from artist_b200 import ops

n = 256
for linear in (True, False):
    case = cases.make_case(n=n, points_per_facet=(4, 4), rays=1, target_pattern=(0, 1))
    kin = case["kin"]
    kin.linear = linear
    if not linear:
        kin.actuator_non_optimizable = kin.actuator_non_optimizable.clone()
        kin.actuator_non_optimizable[:, 0] = 1.0
        kin.actuator_non_optimizable[:, 2] = -10.0
        kin.actuator_non_optimizable[:, 3] = 10.0
    g = torch.Generator().manual_seed(1)
    motor = (20000 + 30000 * torch.rand(n, 2, generator=g)) if linear else (torch.rand(n, 2, generator=g) - 0.3)
    ref = O.motor_positions_to_orientations(kin, motor)
    d = _kin_dev(kin)
    off = O.initial_orientation_offset().reshape(4, 4).to(DEV)
    out = ops.kinematics_orientations(motor.to(DEV), d["rot"], d["trans"], d["opt"] if linear else None, d["positions"],
                                      d["non_opt"], off, linear).cpu()
    same = (out == ref)
    print(f"forward linear={linear}: heliostats with bit-identical orientation {float(same.reshape(n, -1).all(1).float().mean()):.3f}; "
          f"rotation entries identical {float(same[:, :3, :3].float().mean()):.3f}, translation entries {float(same[:, :3, 3].float().mean()):.3f}; "
          f"max diff rot {float((out - ref)[:, :3, :3].abs().max()):.2e} trans {float((out - ref)[:, :3, 3].abs().max()):.2e}")
    ori, mot = O.incident_ray_directions_to_orientations(kin, case["incident"], case["aim"])
    got, gm = ops.kinematics_align_incident(case["incident"].to(DEV), case["aim"].to(DEV), d["rot"], d["trans"],
                                            d["opt"] if linear else None, d["positions"], d["non_opt"], off, linear)
    got, gm = got.cpu(), gm.cpu()
    same = got == ori
    print(f"align   linear={linear}: heliostats with bit-identical orientation {float(same.reshape(n, -1).all(1).float().mean()):.3f}; "
          f"motor positions identical {float((gm == mot).float().mean()):.3f}; rotation entries identical {float(same[:, :3, :3].float().mean()):.3f}; "
          f"max diff rot {float((got - ori)[:, :3, :3].abs().max()):.2e} motor rel {float(((gm - mot).abs() / mot.abs().clamp_min(1)).max()):.2e}")
