"""``artist_b200.compat``: the upstream package's callers (``artist.optim.*``) run unmodified on this package's classes.

CPU part (needs ``/root/reference``, so it runs in the build container only): ``install()`` rebinds every upstream
binding of the hot-path classes - including the ones the optimisers imported with ``from ... import`` and the loader's type
tables - and ``uninstall()`` restores them.  The GPU evidence (upstream ``AimPointOptimizer.optimize`` and the tutorial
flow on a B200, next to upstream's own CUDA-eager run) is ``tools/run_upstream_callers.py`` ->
``profiles/r02_upstream_callers_*.txt``."""
import pytest

from tools.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="upstream package not present (GPU box)")


@pytest.fixture(autouse=True)
def _leave_no_trace():
    """The upstream tree has its own top-level ``tests`` package: take it off ``sys.path`` again (spawned workers of later
    tests inherit the path) and drop the imported upstream modules."""
    import sys

    from tools import ref_import

    path_before = list(sys.path)
    yield
    sys.path[:] = [p for p in path_before if p != ref_import.REFERENCE_ROOT]
    for name in [m for m in sys.modules if m == "artist" or m.startswith("artist.")]:
        del sys.modules[name]


def test_install_rebinds_the_callers_bindings_and_uninstall_restores_them():
    import_reference()
    import artist.optim.aim_point_optimizer as apo
    import artist.optim.kinematics_reconstructor as kr
    import artist.optim.surface_reconstructor as sr
    import artist.util.type_registry as registry

    import artist_b200
    from artist_b200 import compat

    upstream_tracer, upstream_nurbs = apo.HeliostatRayTracer, sr.NURBSSurfaces
    upstream_group = registry.heliostat_group_type_mapping["rigid_body_linear"]
    assert upstream_tracer.__module__.startswith("artist.")
    changed = compat.install()
    try:
        assert compat.install() == []                                        # idempotent
        for mod in (apo, kr, sr):
            assert mod.HeliostatRayTracer is artist_b200.HeliostatRayTracer
            assert mod.Scenario is artist_b200.Scenario
        assert sr.NURBSSurfaces is artist_b200.NURBSSurfaces
        assert registry.heliostat_group_type_mapping["rigid_body_linear"] is artist_b200.HeliostatGroupRigidBody
        assert registry.light_source_type_mapping["sun"] is artist_b200.Sun
        assert "artist.optim.aim_point_optimizer.HeliostatRayTracer" in changed
        # the callers themselves stay upstream code
        assert apo.AimPointOptimizer.__module__ == "artist.optim.aim_point_optimizer"
    finally:
        compat.uninstall()
    assert apo.HeliostatRayTracer is upstream_tracer and sr.NURBSSurfaces is upstream_nurbs
    assert registry.heliostat_group_type_mapping["rigid_body_linear"] is upstream_group


def test_signatures_the_upstream_callers_rely_on():
    """Constructor / method parameters of the replaced classes as the upstream callers pass them (keyword names)."""
    import inspect

    import_reference()
    from artist.field.heliostat_group_rigid_body import HeliostatGroupRigidBody as UpGroup
    from artist.nurbs.surfaces import NURBSSurfaces as UpNurbs
    from artist.raytracing.heliostat_ray_tracer import HeliostatRayTracer as UpTracer

    import artist_b200

    def params(fn):
        return [p for p in inspect.signature(fn).parameters if p != "self"]

    for up, own, methods in (
        (UpTracer, artist_b200.HeliostatRayTracer, ("__init__", "trace_rays", "get_bitmaps_per_target", "get_sampler_indices")),
        (UpNurbs, artist_b200.NURBSSurfaces, ("__init__", "calculate_surface_points_and_normals")),
        (UpGroup, artist_b200.HeliostatGroupRigidBody, ("__init__", "activate_heliostats", "align_surfaces_with_incident_ray_directions",
                                                        "align_surfaces_with_motor_positions")),
    ):
        for m in methods:
            want, got = params(getattr(up, m)), params(getattr(own, m))
            assert got[:len(want)] == want, f"{own.__name__}.{m}: {got} vs upstream {want}"
