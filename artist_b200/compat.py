"""Run the UPSTREAM package's own callers (``artist.optim.*``, tutorials, examples) on top of this package.

``install()`` rebinds, in every loaded ``artist.*`` module, the names that refer to the upstream classes of the
hot path (``HeliostatRayTracer``, ``NURBSSurfaces``, the heliostat-group / kinematics / scenario / tower / sun classes)
to the ``artist_b200`` classes of the same name - by object identity, so ``from artist.raytracing.heliostat_ray_tracer
import HeliostatRayTracer`` inside ``artist/optim/aim_point_optimizer.py:10`` (and every other ``from ... import``
that already ran) resolves to the CUDA implementation, as do class look-up tables such as the loader's
``heliostat_group_type_mapping``.  Nothing upstream is edited; ``uninstall()`` restores every binding.

Typical use (``tools/run_upstream_callers.py``)::

    import artist                      # the unmodified upstream package
    import artist.optim.aim_point_optimizer
    from artist_b200 import compat
    compat.install()
    scenario = artist.scenario.scenario.Scenario.load_scenario_from_hdf5(...)   # now artist_b200's loader
    artist.optim.aim_point_optimizer.AimPointOptimizer(...).optimize(...)      # upstream code, CUDA kernels underneath

The upstream package is never imported by this module unless ``install()`` is called, and nothing else in
``artist_b200`` imports this module.
"""
from __future__ import annotations

import importlib
import sys

# upstream module -> (artist_b200 module, class names defined there)
CLASS_MAP: dict[str, tuple[str, tuple[str, ...]]] = {
    "artist.raytracing.heliostat_ray_tracer": ("artist_b200.raytracing.heliostat_ray_tracer", ("HeliostatRayTracer",)),
    "artist.raytracing.sampling": ("artist_b200.raytracing.sampling", ("DistortionsDataset", "RestrictedDistributedSampler")),
    "artist.nurbs.surfaces": ("artist_b200.nurbs.surfaces", ("NURBSSurfaces",)),
    "artist.field.heliostat_group": ("artist_b200.field.heliostat_group", ("HeliostatGroup",)),
    "artist.field.heliostat_group_rigid_body": ("artist_b200.field.heliostat_group_rigid_body", ("HeliostatGroupRigidBody",)),
    "artist.field.kinematics_rigid_body": ("artist_b200.field.kinematics_rigid_body", ("RigidBody",)),
    "artist.field.heliostat_field": ("artist_b200.field.heliostat_field", ("HeliostatField",)),
    "artist.field.solar_tower": ("artist_b200.field.solar_tower", ("SolarTower",)),
    "artist.field.tower_target_areas_planar": ("artist_b200.field.tower_target_areas", ("TowerTargetAreasPlanar",)),
    "artist.field.tower_target_areas_cylindrical": ("artist_b200.field.tower_target_areas", ("TowerTargetAreasCylindrical",)),
    "artist.scenario.scenario": ("artist_b200.scenario.scenario", ("Scenario",)),
    "artist.scene.sun": ("artist_b200.scene.sun", ("Sun",)),
    "artist.scene.light_source_array": ("artist_b200.scene.light_source_array", ("LightSourceArray",)),
}

_undo: list[tuple[object, object, object]] = []   # (container, key or attribute name, upstream object)


def _pairs() -> dict[int, tuple[object, object]]:
    pairs = {}
    for up_name, (own_name, names) in CLASS_MAP.items():
        try:
            up = importlib.import_module(up_name)
        except ImportError:
            continue
        own = importlib.import_module(own_name)
        for n in names:
            if hasattr(up, n) and hasattr(own, n):
                pairs[id(getattr(up, n))] = (getattr(up, n), getattr(own, n))
    return pairs


def install() -> list[str]:
    """Rebind the upstream classes to this package's; returns ``"module.name"`` of every binding changed."""
    if _undo:
        return []
    if "artist" not in sys.modules:
        raise ImportError("import the upstream `artist` package (and the callers you want to run) before compat.install()")
    pairs = _pairs()
    changed = []
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == "artist" or mod_name.startswith("artist.")):
            continue
        for attr, val in list(vars(mod).items()):
            hit = pairs.get(id(val))
            if hit is not None and hit[0] is val:
                setattr(mod, attr, hit[1])
                _undo.append((mod, attr, val))
                changed.append(f"{mod_name}.{attr}")
            elif isinstance(val, dict):      # class look-up tables (e.g. type name -> group class)
                for k, v in list(val.items()):
                    try:
                        h = pairs.get(id(v))
                    except TypeError:
                        continue
                    if h is not None and h[0] is v:
                        val[k] = h[1]
                        _undo.append((val, k, v))
                        changed.append(f"{mod_name}.{attr}[{k!r}]")
    return changed


def uninstall() -> None:
    while _undo:
        container, key, original = _undo.pop()
        if isinstance(container, dict):
            container[key] = original
        else:
            setattr(container, key, original)
