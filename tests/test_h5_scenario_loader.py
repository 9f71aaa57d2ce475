"""CPU: the built-in HDF5 reader (artist_b200.io.h5lite) and the scenario parser against the fixture the REAL reference
produced from its own scenario files (tests/golden/make_scenario_golden.py): every tensor the reference's loader built
from ``single_heliostat_scenario.h5`` (BASELINE config 1) and ``test_scenario_paint_four_heliostats.h5`` (config 2) must
come out of ``h5_scenario_parser.parse_scenario`` bit for bit."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scenario_golden.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


@pytest.fixture(scope="module")
def files(golden, tmp_path_factory):
    d = tmp_path_factory.mktemp("scenarios")
    out = {}
    for key, entry in golden.items():
        path = d / f"{key}.h5"
        path.write_bytes(entry["file_bytes"].numpy().tobytes())
        out[key] = str(path)
    return out


def test_reader_surface(files):
    from artist_b200.io import h5lite

    with h5lite.File(files["single_heliostat"]) as f:
        assert list(f.keys()) == sorted(f.keys()) and "heliostats" in f and "nope" not in f
        assert f.attrs["version"] == 1.0
        assert len(f["heliostats"]) == 1 and list(f["heliostats"]) == ["heliostat_1"]
        assert f.get("prototypes/surface/facets/facet_9") is None and f.get("x/y", 5) == 5
        with pytest.raises(KeyError):
            f["prototypes/nothing"]
        assert f["lightsources/sun/type"][()] == b"sun"                             # variable-length string -> bytes
        assert f["prototypes/actuator/actuator_2/clockwise_axis_movement"][()] is np.True_   # numpy bool enum
        assert bool(f["prototypes/actuator/actuator_1/clockwise_axis_movement"][()]) is False
        rays = f["lightsources"]["sun"]["number_of_rays"]
        assert rays.shape == () and rays.dtype == np.int64 and int(rays[()]) == 10 and bool(rays)
        cp = f["prototypes/surface/facets/facet_1/control_points"]
        assert cp.shape == (10, 10, 3) and cp.dtype == np.float32 and cp[()].flags.writeable
        assert f["prototypes/surface/facets/facet_1/degrees"][()][1] == 3
        assert float(f["target_areas_planar/planar/plane_e"][()]) == pytest.approx(8.629666667)
    with pytest.raises(h5lite.H5LiteError):
        h5lite.File(files["single_heliostat"], "w")
    bad = files["single_heliostat"] + ".bad"
    open(bad, "wb").write(b"not hdf5 at all")
    with pytest.raises(h5lite.H5LiteError):
        h5lite.File(bad)


@pytest.mark.parametrize("key", ["single_heliostat", "four_heliostats"])
def test_parser_reproduces_the_reference_loader(golden, files, key):
    from artist_b200.io import h5_scenario_parser, h5lite

    ref = golden[key]
    with h5lite.File(files[key]) as f:
        got = h5_scenario_parser.parse_scenario(f)
    assert torch.equal(got["power_plant_position"], ref["power_plant_position"])
    assert got["light_sources"][0]["number_of_rays"] == ref["number_of_rays"]
    assert got["number_of_heliostat_groups"] == len(ref["groups"])
    for name, value in ref["tower"].items():
        if name == "target_name_to_index":
            assert value == {n: i for i, n in enumerate(got["targets"]["planar_names"] + got["targets"]["cyl_names"])}
        elif isinstance(value, list):
            assert got["targets"][name] == value
        else:
            assert torch.equal(got["targets"][name], value), name
    assert len(got["groups"]) == len(ref["groups"])
    for (gkey, g), r in zip(got["groups"].items(), ref["groups"]):     # same groups in the same order
        assert g["names"] == r["names"]
        assert gkey.endswith("linear") == (r["actuator_class"] == "LinearActuators")
        for name in ("positions", "canting", "facet_translations", "nurbs_control_points", "nurbs_degrees",
                     "initial_orientations", "translation_deviations", "rotation_deviations", "actuator_non_optimizable"):
            assert torch.equal(g[name], r[name]), f"{gkey}.{name}"
        assert g["actuator_optimizable"].numel() == r["actuator_optimizable"].numel()
        if r["actuator_optimizable"].numel():
            assert torch.equal(g["actuator_optimizable"], r["actuator_optimizable"])


def test_broken_scenarios_raise_like_the_reference(files):
    from artist_b200.io import h5_scenario_parser, h5lite

    with h5lite.File(files["four_heliostats"]) as f:
        group = f["heliostats/AA31/actuator"]
        with pytest.raises(ValueError, match="wrong amount of actuators"):
            h5_scenario_parser.actuator_parameters("linear", group, 3)
        with pytest.raises(ValueError, match="not yet implemented"):
            h5_scenario_parser.actuator_parameters("hydraulic", group, 2)
        with pytest.raises(ValueError, match="not yet implemented"):
            h5_scenario_parser.kinematics_deviations("new_kinematics", f["heliostats/AA31/kinematics"])
