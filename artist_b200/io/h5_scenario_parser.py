"""ARTIST scenario file (HDF5) -> plain CPU tensors, grouped the way the reference groups heliostats.

Restates what ``Scenario.load_scenario_from_hdf5`` (``artist/scenario/scenario.py:105-259``), the helpers of
``artist/io/h5_scenario_parser.py`` (surface config :12-74, rigid-body deviations :131-333, linear actuators :398-571
incl. the south->up initial-angle fix :540-569, ideal actuators :574-640), ``HeliostatField.from_hdf5``
(``artist/field/heliostat_field.py:81-435``), ``TowerTargetAreas*.from_hdf5`` and ``Sun.from_hdf5`` read from the file:
same keys, same defaults (missing deviation / actuator parameters are 0), same prototype fall-backs, same group key
``"<kinematics type>_<actuator type>"`` in first-appearance order, same errors.  Works with ``h5py.File`` objects as
well as with :mod:`artist_b200.io.h5lite` (only ``[]``, ``.keys()``, ``.get()``, ``[()]``, ``.attrs`` are used).
Everything here is host-side bookkeeping on a few KB of parameters; the surfaces are evaluated later on the GPU.
"""
from __future__ import annotations

import math

import numpy as np
import torch

TRANSLATION_DEVIATIONS = ("first_joint_translation_e", "first_joint_translation_n", "first_joint_translation_u",
                          "second_joint_translation_e", "second_joint_translation_n", "second_joint_translation_u",
                          "concentrator_translation_e", "concentrator_translation_n", "concentrator_translation_u")
ROTATION_DEVIATIONS = ("first_joint_tilt_n", "first_joint_tilt_u", "second_joint_tilt_e", "second_joint_tilt_n")
LINEAR_ACTUATOR_INT, IDEAL_ACTUATOR_INT = 0, 1


def _f32(value) -> torch.Tensor:
    return torch.tensor(np.asarray(value), dtype=torch.float)


def _text(dataset) -> str:
    value = dataset[()]
    return value.decode("utf-8") if isinstance(value, bytes) else str(value)


def _scalar_or_zero(group, path: str) -> torch.Tensor:
    node = group.get(path)
    return _f32(node[()]) if node is not None else torch.tensor(0.0)


def surface_config(facets) -> dict:
    """``facets`` group -> control points ``[F,cu,cv,3]``, degrees ``[2]`` (of the last facet, as the reference
    ends up with), canting ``[F,2,4]``, facet translations ``[F,4]``."""
    names = list(facets.keys())
    return dict(
        control_points=torch.stack([_f32(facets[f]["control_points"][()]) for f in names]),
        degrees=torch.tensor([int(facets[names[-1]]["degrees"][()][0]), int(facets[names[-1]]["degrees"][()][1])],
                             dtype=torch.int32),
        canting=torch.stack([_f32(facets[f]["canting"][()]) for f in names]),
        facet_translations=torch.stack([_f32(facets[f]["position"][()]) for f in names]),
    )


def kinematics_deviations(kinematics_type: str, kinematics_config) -> tuple[torch.Tensor, torch.Tensor, int]:
    if kinematics_type != "rigid_body":
        raise ValueError(f"The kinematics type: {kinematics_type} is not yet implemented!")
    translation = torch.stack([_scalar_or_zero(kinematics_config, f"deviations/{k}") for k in TRANSLATION_DEVIATIONS])
    rotation = torch.stack([_scalar_or_zero(kinematics_config, f"deviations/{k}") for k in ROTATION_DEVIATIONS])
    return translation, rotation, 2


def actuator_parameters(actuator_type: str, actuator_config, number_of_actuators: int) -> tuple[torch.Tensor, torch.Tensor]:
    """-> ``(non_optimizable [7 | 4, A], optimizable [2, A] | empty)`` in the reference's row order: type, clockwise,
    min, max motor position (, increment, offset, pivot radius) / initial angle, initial stroke length."""
    if actuator_type not in ("linear", "ideal"):
        raise ValueError(f"The actuator type: {actuator_type} is not yet implemented!")
    keys = list(actuator_config.keys())
    if len(keys) != number_of_actuators:
        raise ValueError("This scenario file contains the wrong amount of actuators for this heliostat and its kinematics "
                         f"type. Expected {number_of_actuators} actuators, found {len(keys)} actuator(s).")
    linear = actuator_type == "linear"
    non_opt = torch.zeros(7 if linear else 4, number_of_actuators)
    opt = torch.zeros(2, number_of_actuators) if linear else torch.tensor([])
    for j, key in enumerate(keys):
        cfg = actuator_config[key]
        limits = cfg["min_max_motor_positions"][()]
        non_opt[0, j] = LINEAR_ACTUATOR_INT if linear else IDEAL_ACTUATOR_INT
        non_opt[1, j] = 1.0 if bool(cfg["clockwise_axis_movement"][()]) else 0.0
        non_opt[2, j] = float(limits[0])
        non_opt[3, j] = float(limits[1])
        if linear:
            non_opt[4, j] = _scalar_or_zero(cfg, "parameters/increment")
            non_opt[5, j] = _scalar_or_zero(cfg, "parameters/offset")
            non_opt[6, j] = _scalar_or_zero(cfg, "parameters/pivot_radius")
            opt[0, j] = _scalar_or_zero(cfg, "parameters/initial_angle")
            opt[1, j] = _scalar_or_zero(cfg, "parameters/initial_stroke_length")
    if linear:
        # surfaces are sampled facing up, the kinematics' standard orientation is south: the rotation south -> up
        # (axis -e, angle pi/2) projected on actuator one's axis (east) is added to its initial angle (:540-569)
        south, up = torch.tensor([0.0, -1.0, 0.0]), torch.tensor([0.0, 0.0, 1.0])
        axis = torch.nn.functional.normalize(torch.linalg.cross(south, up), dim=0)
        angle = torch.arccos(torch.clamp(torch.dot(south, up), -1.0, 1.0))
        opt[0, 0] += axis[0] * angle
    return non_opt, opt


def parse_targets(f) -> dict:
    planar, cyl = f["target_areas_planar"], f["target_areas_cylindrical"]
    pn, cn = sorted(planar.keys()), sorted(cyl.keys())
    vec4 = lambda d: _f32(d[()]).reshape(-1)[:4]
    return dict(
        planar_names=pn,
        planar_centers=torch.stack([vec4(planar[k]["position_center"]) for k in pn]) if pn else torch.zeros(0, 4),
        planar_normals=torch.stack([vec4(planar[k]["normal_vector"]) for k in pn]) if pn else torch.zeros(0, 4),
        planar_dimensions=torch.tensor([[float(planar[k]["plane_e"][()]), float(planar[k]["plane_u"][()])] for k in pn],
                                       dtype=torch.float).reshape(-1, 2),
        cyl_names=cn,
        cyl_centers=torch.stack([vec4(cyl[k]["cylinder_center"]) for k in cn]) if cn else torch.zeros(0, 4),
        cyl_axes=torch.stack([vec4(cyl[k]["cylinder_axis"]) for k in cn]) if cn else torch.zeros(0, 4),
        cyl_normals=torch.stack([vec4(cyl[k]["cylinder_normal"]) for k in cn]) if cn else torch.zeros(0, 4),
        cyl_radii=torch.stack([_f32(cyl[k]["cylinder_radius"][()]) for k in cn]) if cn else torch.zeros(0),
        cyl_heights=torch.stack([_f32(cyl[k]["cylinder_height"][()]) for k in cn]) if cn else torch.zeros(0),
        cyl_opening_angles=torch.stack([_f32(cyl[k]["cylinder_opening_angle"][()]) for k in cn]) if cn else torch.zeros(0),
    )


def parse_light_sources(f) -> list[dict]:
    out = []
    for name in sorted(f["lightsources"].keys()):
        cfg = f["lightsources"][name]
        kind = _text(cfg["type"])
        if kind != "sun":
            raise KeyError(f"Currently the selected light source: {kind} is not supported.")
        dp = cfg["distribution_parameters"]
        params = {"distribution_type": _text(dp["distribution_type"])}
        for key in ("mean", "covariance"):
            if key in dp.keys():
                params[key] = float(dp[key][()])
        out.append(dict(name=name, number_of_rays=int(cfg["number_of_rays"][()]), distribution_parameters=params))
    return out


def _prototype(f) -> dict:
    proto = f["prototypes"]
    kin_type = _text(proto["kinematics"]["type"])
    tdev, rdev, n_act = kinematics_deviations(kin_type, proto["kinematics"])
    keys = list(proto["actuator"].keys())
    types = [_text(proto["actuator"][k]["type"]) for k in keys]
    if not types:
        raise ValueError("Prototype actuator type list is empty.")
    if len(set(types)) > 1:
        raise ValueError("Prototype actuators must all have the same type.")
    non_opt, opt = actuator_parameters(types[0], proto["actuator"], n_act)
    return dict(surface=surface_config(proto["surface"]["facets"]), kinematics_type=kin_type,
                initial_orientation=_f32(proto["kinematics"]["initial_orientation"][()]), translation_deviations=tdev,
                rotation_deviations=rdev, number_of_actuators=n_act, actuator_type=types[0], actuator_non_optimizable=non_opt,
                actuator_optimizable=opt)


def parse_scenario(f) -> dict:
    """-> ``dict(version, power_plant_position, targets, light_sources, groups)``; ``groups`` maps the reference's
    group key to stacked per-heliostat tensors (the layout ``synthetic_field_tensors`` uses)."""
    proto = _prototype(f)
    groups: dict[str, dict] = {}
    for name in f["heliostats"].keys():
        h = f["heliostats"][name]
        members = set(h.keys())
        surface = surface_config(h["surface"]["facets"]) if "surface" in members else proto["surface"]
        if "kinematics" in members:
            kin_type = _text(h["kinematics"]["type"])
            orientation = _f32(h["kinematics"]["initial_orientation"][()])
            tdev, rdev, n_act = kinematics_deviations(kin_type, h["kinematics"])
        else:
            kin_type, orientation = proto["kinematics_type"], proto["initial_orientation"]
            tdev, rdev, n_act = proto["translation_deviations"], proto["rotation_deviations"], proto["number_of_actuators"]
        if "actuator" in members:
            types = [_text(h["actuator"][k]["type"]) for k in h["actuator"].keys()]
            if len(set(types)) > 1:
                raise ValueError("When using the rigid body kinematics, all actuators for a given heliostat must have the "
                                 "same type.")
            act_type = types[0]
            non_opt, opt = actuator_parameters(act_type, h["actuator"], n_act)
        else:
            act_type, non_opt, opt = proto["actuator_type"], proto["actuator_non_optimizable"], proto["actuator_optimizable"]
        g = groups.setdefault(f"{kin_type}_{act_type}", dict(kinematics_type=kin_type, actuator_type=act_type, names=[], rows=[]))
        g["names"].append(name)
        g["rows"].append(dict(positions=_f32(h["position"][()]), nurbs_control_points=surface["control_points"],
                              canting=surface["canting"], facet_translations=surface["facet_translations"],
                              initial_orientations=orientation, translation_deviations=tdev, rotation_deviations=rdev,
                              actuator_non_optimizable=non_opt, actuator_optimizable=opt, degrees=surface["degrees"]))
    for g in groups.values():
        rows = g.pop("rows")
        for key in rows[0]:
            if key != "degrees":
                g[key] = torch.stack([r[key] for r in rows])
        g["nurbs_degrees"] = rows[-1]["degrees"]
        # kernel layout: per heliostat [rows, actuators] -> [N, rows, 2] (already the case after stacking)
    return dict(version=f.attrs["version"] if "version" in f.attrs else None,
                power_plant_position=torch.tensor(np.asarray(f["power_plant"]["position"][()]), dtype=torch.float64),
                number_of_heliostat_groups=int(f["number_of_heliostat_groups"][()]) if "number_of_heliostat_groups" in f.keys() else None,
                targets=parse_targets(f), light_sources=parse_light_sources(f), groups=groups)


assert math.isclose(float(torch.arccos(torch.tensor(0.0))), math.pi / 2, rel_tol=1e-6)
