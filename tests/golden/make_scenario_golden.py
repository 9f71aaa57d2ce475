"""Generate ``tests/golden/scenario_golden.pt``: the REAL reference (imported from ``/root/reference``, build container
only) loading its own tutorial / test scenario FILES and tracing them on the CPU - BASELINE.json configs[0] (tutorial 01:
``single_heliostat_scenario.h5``, four sun directions, 256x256 bitmap) and configs[1] (``test_scenario_paint_four_
heliostats.h5``: two heliostat groups, ideal + linear actuators, 20x20 control points per facet).

``h5py`` does not exist in this image, so the reference reads the files through ``artist_b200.io.h5lite`` (installed as
its ``h5py.File``) - which makes this script also the end-to-end check of that reader: the reference's own loaders must
accept what it returns.  The fixture stores, per scenario: the scenario file's bytes (inputs for the loader tests - the
GPU box has no /root/reference), every tensor the reference's loader produced (pins ``h5_scenario_parser``), and the
reference's outputs (aligned surfaces, motor positions, flux bitmaps as sparse tensors, factors).
Re-run:  python tests/golden/make_scenario_golden.py
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from artist_b200.io import h5lite  # noqa: E402
from tools.ref_import import REFERENCE_ROOT, import_reference  # noqa: E402

CPU = torch.device("cpu")
SCENARIOS = {
    "single_heliostat": "tutorials/data/scenarios/single_heliostat_scenario.h5",
    "four_heliostats": "tests/data/scenarios/test_scenario_paint_four_heliostats.h5",
}


def group_tensors(group) -> dict:
    kin, act = group.kinematics, group.kinematics.actuators
    return dict(names=list(group.names), positions=group.positions.clone(), canting=group.canting.clone(),
                facet_translations=group.facet_translations.clone(), nurbs_control_points=group.nurbs_control_points.clone(),
                nurbs_degrees=group.nurbs_degrees.clone(), initial_orientations=group.initial_orientations.clone(),
                translation_deviations=kin.translation_deviation_parameters.clone(),
                rotation_deviations=kin.rotation_deviation_parameters.clone(),
                actuator_non_optimizable=act.non_optimizable_parameters.clone(),
                actuator_optimizable=act.optimizable_parameters.clone(),
                surface_points_sample=group.surface_points[:, ::SAMPLE].clone(),
                surface_normals_sample=group.surface_normals[:, ::SAMPLE].clone(),
                group_class=type(group).__name__, actuator_class=type(act).__name__)


def tower_tensors(tower) -> dict:
    planar, cyl = tower.target_areas[0], tower.target_areas[1]
    return dict(planar_names=list(planar.names), planar_centers=planar.centers.clone(), planar_normals=planar.normals.clone(),
                planar_dimensions=planar.dimensions.clone(), cyl_names=list(cyl.names), cyl_centers=cyl.centers.clone(),
                cyl_normals=cyl.normals.clone(), cyl_axes=cyl.axes.clone(), cyl_radii=cyl.radii.clone(),
                cyl_heights=cyl.heights.clone(), cyl_opening_angles=cyl.opening_angles.clone(),
                target_name_to_index=dict(tower.target_name_to_index))


SAMPLE = 53   # surfaces are stored as every 53rd point (the full tensors would make the fixture 16 MB)


def pack_sparse(x: torch.Tensor) -> dict:
    """Dense bitmap stack -> flat int32 indices + float32 values of its non-zero pixels (exact, a few KB)."""
    flat = x.reshape(-1)
    idx = flat.nonzero().reshape(-1)
    return dict(shape=tuple(x.shape), idx=idx.to(torch.int32), val=flat[idx].clone())


def trace(scenario, group, inc, mask, tidx, **tracer_kwargs):
    from artist.raytracing.heliostat_ray_tracer import HeliostatRayTracer

    group.activate_heliostats(active_heliostats_mask=mask, device=CPU)
    aim = scenario.solar_tower.get_centers_of_target_areas(target_area_indices=tidx, device=CPU)
    group.align_surfaces_with_incident_ray_directions(aim_points=aim, incident_ray_directions=inc,
                                                      active_heliostats_mask=mask, device=CPU)
    tracer = HeliostatRayTracer(scenario=scenario, heliostat_group=group, **tracer_kwargs)
    flux, ic, ot, bl = tracer.trace_rays(incident_ray_directions=inc, active_heliostats_mask=mask,
                                         target_area_indices=tidx, device=CPU)
    per_target = tracer.get_bitmaps_per_target(flux, tidx, device=CPU)
    du, de = tracer.distortions_dataset.distortions_u, tracer.distortions_dataset.distortions_e
    checksum = dict(shape=tuple(du.shape), sum_u=float(du.double().sum()), sum_e=float(de.double().sum()),
                    head_u=du.flatten()[:8].clone(), tail_e=de.flatten()[-8:].clone())
    return dict(incident=inc.clone(), mask=mask.clone(), target_idx=tidx.clone(), aim=aim.clone(),
                motor_positions=group.kinematics.active_motor_positions.clone(),
                aligned_points_sample=group.active_surface_points[:, ::SAMPLE].clone(),
                aligned_normals_sample=group.active_surface_normals[:, ::SAMPLE].clone(),
                flux=pack_sparse(flux), per_target=pack_sparse(per_target), intercept=ic.clone(),
                on_target=ot.clone(), blocking=bl.clone(), ray_magnitude=float(tracer.ray_magnitude),
                distortions=checksum, resolution=tuple(int(v) for v in tracer.bitmap_resolution))


def main() -> None:
    import_reference()
    sys.modules["h5py"].File = h5lite.File      # the reference's loaders now read through the reader under test
    from artist.field.heliostat_field import HeliostatField  # noqa: F401  (import order avoids the reference's import cycle)
    from artist.scenario.scenario import Scenario

    out = {}
    for key, rel in SCENARIOS.items():
        path = os.path.join(REFERENCE_ROOT, rel)
        with open(path, "rb") as fh:
            file_bytes = fh.read()
        with h5lite.File(path) as f:
            scenario = Scenario.load_scenario_from_hdf5(scenario_file=f, device=CPU)
        groups = scenario.heliostat_field.heliostat_groups
        entry = dict(source=rel, file_bytes=torch.frombuffer(bytearray(file_bytes), dtype=torch.uint8).clone(),
                     power_plant_position=scenario.power_plant_position.clone(), tower=tower_tensors(scenario.solar_tower),
                     number_of_rays=int(scenario.light_sources.light_source_list[0].number_of_rays),
                     groups=[group_tensors(g) for g in groups], traces=[])
        if key == "single_heliostat":           # tutorials/01_single_heliostat_raytracing_tutorial.py
            mask, tidx = torch.tensor([1], dtype=torch.int32), torch.tensor([0])
            for name, d in (("south", [0.0, 1.0, 0.0, 0.0]), ("east", [-1.0, 0.0, 0.0, 0.0]), ("west", [1.0, 0.0, 0.0, 0.0]),
                            ("above", [0.0, 0.0, -1.0, 0.0])):
                t = trace(scenario, groups[0], torch.tensor([d]), mask, tidx)
                t["name"], t["group"] = name, 0
                entry["traces"].append(t)
        else:                                   # every group, default mapping, towards the multi-focus tower (index 0) and the receiver
            for gi, g in enumerate(groups):
                for target in (0, 3):
                    mask, tidx, inc = scenario.index_mapping(heliostat_group=g, single_incident_ray_direction=torch.tensor(
                        [0.0, 0.9805806756909201, -0.19611613513818402, 0.0]), single_target_area_index=target, device=CPU)
                    t = trace(scenario, g, inc, mask, tidx, bitmap_resolution=torch.tensor([64, 64]))
                    t["name"], t["group"] = f"group{gi}_target{target}", gi
                    entry["traces"].append(t)
        for t in entry["traces"]:
            f = torch.zeros(t["flux"]["shape"]).reshape(-1)
            f[t["flux"]["idx"].long()] = t["flux"]["val"]
            print(f"{key:18s} {t['name']:16s} flux sum {f.sum():10.3f} peak {f.max():8.4f} nnz {int((f != 0).sum()):6d} "
                  f"intercept {t['intercept'].tolist()} blocking {t['blocking'].tolist()}")
        out[key] = entry
    dst = os.path.join(ROOT, "tests", "golden", "scenario_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
