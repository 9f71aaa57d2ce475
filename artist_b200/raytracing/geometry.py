"""The reference's exported ray-geometry functions (``artist/raytracing/geometry.py``) on its materialised per-ray
tensors, as stand-alone CUDA kernels behind the C ABI (``csrc/geometry.cu``).

``HeliostatRayTracer.trace_rays`` does NOT call these - its fused kernels keep every per-ray quantity in registers.
They exist so that code written against the reference's step-wise API keeps working and so that each step can be
checked on its own against the reference's unit tests.  Same signatures, shapes and value semantics as the reference;
forward only (the differentiable path is ``trace_rays``)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib, ops
from ..scene.rays import Rays


def _warn_if_grad(what: str, *tensors) -> None:
    """The reference's step functions are differentiable; these stand-alone forms are forward-only (API parity and
    step-wise checks - the differentiable path is the fused ``trace_rays``).  Say so once instead of silently dropping
    the graph."""
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors):
        ops.warn_no_graph(what)


def reflect(incident_ray_directions: torch.Tensor, reflection_surface_normals: torch.Tensor) -> torch.Tensor:
    """``r = i - 2 (i . n) n`` (``geometry.py:11-41``): incident ``[N,1,4]``, normals ``[N,P,4]`` -> ``[N,P,4]``."""
    _warn_if_grad("reflect", incident_ray_directions, reflection_surface_normals)
    normals = ops._f32(reflection_surface_normals.detach(), "reflection_surface_normals")
    n, p, _ = normals.shape
    incident = ops._f32(incident_ray_directions.detach().expand(n, 1, 4).reshape(n, 4), "incident_ray_directions")
    out = torch.empty_like(normals)
    _lib.call("ab200_reflect", ops._p(incident), ops._p(normals), n, p, ops._p(out), ops._stream())
    return out


def _intersections(rays: Rays, points_at_ray_origins, targets: ops.TargetTensors, target_area_indices, bitmap_resolution,
                   cylindrical: bool):
    _warn_if_grad("line_cylinder_intersections" if cylindrical else "line_plane_intersections", rays.ray_directions,
                  rays.ray_magnitudes, points_at_ray_origins)
    dirs = ops._f32(rays.ray_directions.detach(), "ray_directions")
    mags = ops._f32(rays.ray_magnitudes.detach(), "ray_magnitudes")
    origins = ops._f32(points_at_ray_origins.detach(), "points_at_ray_origins")
    n, r, p, _ = dirs.shape
    tidx = None if target_area_indices is None else ops._i32(target_area_indices, "target_area_indices")
    res_e, res_u = int(bitmap_resolution[0]), int(bitmap_resolution[1])
    out = [torch.empty(n, r, p, device=dirs.device) for _ in range(4)]
    ts = targets.struct()
    _lib.call("ab200_line_intersections", ops._p(dirs), ops._p(mags), ops._p(origins), C.byref(ts), ops._p(tidx),
              1 if cylindrical else 0, n, r, p, res_e, res_u, *(ops._p(o) for o in out), ops._stream())
    return tuple(out)


def _target_tensors(planar=None, cylindrical=None, device=None) -> ops.TargetTensors:
    f = lambda x, shape: torch.as_tensor(x, dtype=torch.float32, device=device).reshape(shape).contiguous()
    z4, z1, z2 = torch.zeros(0, 4, device=device), torch.zeros(0, device=device), torch.zeros(0, 2, device=device)
    if planar is not None:
        return ops.TargetTensors(f(planar.centers, (-1, 4)), f(planar.normals, (-1, 4)), f(planar.dimensions, (-1, 2)),
                                 z4, z4, z4, z1, z1, z1)
    return ops.TargetTensors(z4, z4, z2, f(cylindrical.centers, (-1, 4)), f(cylindrical.normals, (-1, 4)),
                             f(cylindrical.axes, (-1, 4)), f(cylindrical.radii, (-1,)), f(cylindrical.heights, (-1,)),
                             f(cylindrical.opening_angles, (-1,)))


def line_plane_intersections(rays: Rays, points_at_ray_origins: torch.Tensor, target_areas,
                             target_area_indices: torch.Tensor | None = None,
                             bitmap_resolution: torch.Tensor = torch.tensor([256, 256]), device=None):
    """``geometry.py:44-204`` -> ``(bitmap_coordinates_e, bitmap_coordinates_u, intersection_distances,
    absolute_intensities)``, each ``[N,R,P]``; invalid rays are zero (e-coordinate: ``E-1`` after the flip)."""
    dev = rays.ray_directions.device
    return _intersections(rays, points_at_ray_origins, _target_tensors(planar=target_areas, device=dev), target_area_indices,
                          bitmap_resolution, cylindrical=False)


def line_cylinder_intersections(rays: Rays, points_at_ray_origins: torch.Tensor, target_areas,
                                target_area_indices: torch.Tensor | None = None,
                                bitmap_resolution: torch.Tensor = torch.tensor([256, 256]), device=None):
    """``geometry.py:207-445``; ``target_area_indices`` count within the cylindrical areas."""
    dev = rays.ray_directions.device
    return _intersections(rays, points_at_ray_origins, _target_tensors(cylindrical=target_areas, device=dev),
                          target_area_indices, bitmap_resolution, cylindrical=True)


def bilinear_splatting(bitmap_intersections_e: torch.Tensor, bitmap_intersections_u: torch.Tensor,
                       absolute_intensities: torch.Tensor, bitmap_resolution) -> torch.Tensor:
    """``heliostat_ray_tracer.py:610-778``: three ``[N,...]`` tensors -> flux bitmaps ``[N,U,E]``."""
    n = absolute_intensities.shape[0]
    _warn_if_grad("bilinear_splatting", bitmap_intersections_e, bitmap_intersections_u, absolute_intensities)
    be = ops._f32(bitmap_intersections_e.detach().reshape(n, -1), "bitmap_intersections_e")
    bu = ops._f32(bitmap_intersections_u.detach().reshape(n, -1), "bitmap_intersections_u")
    v = ops._f32(absolute_intensities.detach().reshape(n, -1), "absolute_intensities")
    res_e, res_u = int(bitmap_resolution[0]), int(bitmap_resolution[1])
    out = torch.empty(n, res_u, res_e, device=v.device)
    _lib.call("ab200_bilinear_splatting", ops._p(be), ops._p(bu), ops._p(v), n, be.shape[1], res_e, res_u, ops._p(out),
              ops._stream())
    return out
