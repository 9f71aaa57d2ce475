"""CPU: the C-ABI library loads and exports every symbol include/artist_b200.h declares; host-side logic
(index mapping, sampler, activation, distributed setup helpers) works without a GPU; the product path
refuses CPU tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "artist_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ab200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from artist_b200 import _build, _lib

    lib = ctypes.CDLL(_build.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTS) == set(declared), "ctypes binding and header disagree"
    assert _lib.lib().ab200_abi_version() == _lib.ABI_VERSION == 4
    assert _lib.lib().ab200_error_string(-1) == b"invalid argument"


def test_struct_sizes_match_the_header():
    """sizeof() as compiled by gcc from the header == ctypes layout (catches binding drift)."""
    import subprocess
    import tempfile

    from artist_b200 import _lib

    src = '#include <stdio.h>\n#include "artist_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ab200_targets), sizeof(ab200_blockers), sizeof(ab200_trace_args), sizeof(ab200_trace_bwd_args), sizeof(ab200_nurbs_args), sizeof(ab200_kinematics_args), sizeof(ab200_host_trace_args));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(t) for t in (_lib.Targets, _lib.Blockers, _lib.TraceArgs, _lib.TraceBwdArgs, _lib.NurbsArgs,
                                       _lib.KinematicsArgs, _lib.HostTraceArgs)]
    assert sizes == want


def test_argument_validation_without_gpu():
    from artist_b200 import _lib

    lib = _lib.lib()
    assert lib.ab200_trace_fwd(None, None) == -1
    a = _lib.TraceArgs()
    a.abi_version = 99
    assert lib.ab200_trace_fwd(ctypes.byref(a), None) == -1
    assert b"abi_version" in lib.ab200_last_error_detail()
    n = _lib.NurbsArgs()
    n.abi_version, n.n_surfaces, n.n_facets, n.n_eval, n.degree_u, n.degree_v = _lib.ABI_VERSION, 1, 1, 1, 5, 3
    assert lib.ab200_nurbs_fwd(ctypes.byref(n), None) == -1
    assert b"degree" in lib.ab200_last_error_detail()


def test_no_cpu_fallback():
    from artist_b200 import _lib, ops

    tg = ops.TargetTensors(*[torch.zeros(1, 4)] * 2, torch.ones(1, 2), *[torch.zeros(0, 4)] * 3, *[torch.zeros(0)] * 3)
    with pytest.raises(_lib.Ab200Error, match="CUDA tensor"):
        ops.trace(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4), torch.zeros(1, 4), torch.zeros(1, 2, 4, 2),
                  torch.zeros(1, dtype=torch.int32), tg, ops.TraceOptions())


def test_no_oracle_import_in_product():
    """The package must never import the oracle (or tests/tools)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "artist_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests|tools)\b", text, flags=re.M), os.path.join(dirpath, f)


def _cpu_scenario(n=3):
    """Host objects on the CPU (containers and index logic only; no kernels are called)."""
    from artist_b200.field import (HeliostatField, HeliostatGroupRigidBody, SolarTower, TowerTargetAreasCylindrical,
                                   TowerTargetAreasPlanar)
    from artist_b200.scenario import Scenario, synthetic_field_tensors
    from artist_b200.scene import LightSourceArray, Sun

    ft = synthetic_field_tensors(n, control_points=(6, 6))
    pts = torch.zeros(n, 16, 4)
    group = HeliostatGroupRigidBody(
        names=ft["names"], positions=ft["positions"], surface_points=pts, surface_normals=pts.clone(), canting=ft["canting"],
        facet_translations=ft["facet_translations"], initial_orientations=ft["initial_orientations"],
        nurbs_control_points=ft["nurbs_control_points"], nurbs_degrees=ft["nurbs_degrees"],
        kinematics_translation_deviation_parameters=ft["translation_deviations"],
        kinematics_rotation_deviation_parameters=ft["rotation_deviations"],
        actuator_parameters_non_optimizable=ft["actuator_non_optimizable"],
        actuator_parameters_optimizable=ft["actuator_optimizable"], device=torch.device("cpu"))
    planar = TowerTargetAreasPlanar(ft["planar_names"], ft["planar_centers"], ft["planar_normals"], ft["planar_dimensions"])
    cyl = TowerTargetAreasCylindrical(ft["cyl_names"], ft["cyl_centers"], ft["cyl_normals"], ft["cyl_axes"], ft["cyl_radii"],
                                      ft["cyl_heights"], ft["cyl_opening_angles"])
    scenario = Scenario(torch.zeros(3, dtype=torch.float64), SolarTower([planar, cyl], device=torch.device("cpu")),
                        LightSourceArray([Sun(4, device=torch.device("cpu"))]), HeliostatField([group], device=torch.device("cpu")))
    return scenario, group


def test_index_mapping_and_activation():
    scenario, group = _cpu_scenario(3)
    cpu = torch.device("cpu")
    mask, tidx, inc = scenario.index_mapping(group, device=cpu)
    assert mask.tolist() == [1, 1, 1] and mask.dtype == torch.int32 and tidx.tolist() == [0, 0, 0]
    assert inc.shape == (3, 4) and inc[0].tolist() == [0.0, 1.0, 0.0, 0.0]
    d = torch.tensor([0.0, 0.6, -0.8, 0.0])
    mapping = [("H00002", "receiver", d), ("H00000", "receiver_plane", d), ("H00002", "receiver_plane", d), ("nope", "x", d)]
    mask, tidx, inc = scenario.index_mapping(group, string_mapping=mapping, device=cpu)
    assert mask.tolist() == [1, 0, 2] and tidx.tolist() == [0, 1, 0]
    with pytest.raises(ValueError, match="Invalid target"):
        scenario.index_mapping(group, string_mapping=[("H00000", "missing", d)], device=cpu)
    with pytest.raises(ValueError, match="invalid"):
        scenario.index_mapping(group, single_incident_ray_direction=torch.tensor([0.0, 2.0, 0.0, 0.0]), device=cpu)
    with pytest.raises(ValueError, match="target area index"):
        scenario.index_mapping(group, single_target_area_index=7, device=cpu)
    group.activate_heliostats(mask, device=cpu)
    assert group.number_of_active_heliostats == 3
    assert group.active_surface_points.shape[0] == 3 and group.kinematics.active_heliostat_positions.shape[0] == 3
    assert torch.equal(group.kinematics.active_heliostat_positions[1], group.positions[2])
    aim = scenario.solar_tower.get_centers_of_target_areas(tidx)
    assert torch.allclose(aim[0], torch.tensor([0.0, 0.0, 50.0, 1.0]))
    assert torch.allclose(aim[1, :3], torch.tensor([0.0, -3.76, 56.7]) + 4.14 * torch.tensor([0.0, 0.9063, -0.4226]))
    assert scenario.solar_tower.target_name_to_index == {"receiver_plane": 0, "receiver": 1}


def test_group_distribution_matches_reference_contract():
    from artist_b200.util.env import distribute_groups_among_ranks

    assert distribute_groups_among_ranks(1, 3) == ({0: [0, 1, 2]}, False)
    assert distribute_groups_among_ranks(2, 3) == ({0: [0, 2], 1: [1]}, False)
    assert distribute_groups_among_ranks(4, 2) == ({0: [0], 1: [1], 2: [0], 3: [1]}, True)


def test_sun_matches_reference_rng_contract():
    from artist_b200.scene import Sun
    from oracle import artist_oracle as O

    du, de = Sun(3, device=torch.device("cpu")).get_distortions(5, 2, random_seed=7)
    ou, oe = O.sun_distortions(3, 5, 2, 7)
    assert torch.equal(du, ou) and torch.equal(de, oe)
    from artist_b200.ops import pack_distortions

    packed = pack_distortions(du, de)
    assert packed.data_ptr() == du.data_ptr() and packed.shape == (2, 3, 5, 2), "views of one buffer: no copy"
    assert torch.equal(pack_distortions(du.clone(), de.clone())[..., 1], de)
