"""CPU oracle for the ARTIST heliostat ray-tracing hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (torch CPU tensors, fp32, eager ops + autograd)
of the algorithm behind ``HeliostatRayTracer.trace_rays`` and the two steps that feed it
(NURBS surface evaluation, rigid-body alignment) in ARTIST v2.0.0.  It works on plain
tensors, not on scenario objects.  Every function cites the reference ``file:line`` it
follows (paths relative to the upstream repository root).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module, and only as the checker / the CPU arm that is timed *beside* the product.
Nothing under ``artist_b200/`` imports it: the product path is the CUDA library and fails
loudly without it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the real reference (imported
from ``/root/reference`` in the build container) and this oracle on identical inputs; the
committed fixtures under ``tests/golden/`` hold the reference's outputs, and
``tests/test_oracle_golden.py`` checks the oracle against them (bit-exact for pixel
coordinates / bitmaps on CPU, 1e-6-level for NURBS and kinematics).  The reference's own
inline known-answer tests for the steps (reflect, line-plane, line-cylinder,
rotate_distortions, NURBS forward, kinematics orientation matrices, sampler index lists)
are restated in ``tests/test_oracle_kat.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------


@dataclass
class Targets:
    """Target-area SoA tensors (``artist/field/tower_target_areas_{planar,cylindrical}.py``).

    Global target index = planar areas first, then cylindrical ones
    (``artist/field/solar_tower.py:80-90``).
    """

    planar_centers: torch.Tensor = field(default_factory=lambda: torch.zeros(0, 4))
    planar_normals: torch.Tensor = field(default_factory=lambda: torch.zeros(0, 4))
    planar_dimensions: torch.Tensor = field(default_factory=lambda: torch.zeros(0, 2))
    cyl_centers: torch.Tensor = field(default_factory=lambda: torch.zeros(0, 4))
    cyl_normals: torch.Tensor = field(default_factory=lambda: torch.zeros(0, 4))
    cyl_axes: torch.Tensor = field(default_factory=lambda: torch.zeros(0, 4))
    cyl_radii: torch.Tensor = field(default_factory=lambda: torch.zeros(0))
    cyl_heights: torch.Tensor = field(default_factory=lambda: torch.zeros(0))
    cyl_opening_angles: torch.Tensor = field(default_factory=lambda: torch.zeros(0))

    @property
    def n_planar(self) -> int:
        return int(self.planar_centers.shape[0])

    @property
    def n_cyl(self) -> int:
        return int(self.cyl_centers.shape[0])

    @property
    def n_total(self) -> int:
        return self.n_planar + self.n_cyl


def targets_from_field_tensors(ft: dict) -> Targets:
    """Targets from a ``synthetic_field_tensors`` dictionary."""
    return Targets(ft["planar_centers"], ft["planar_normals"], ft["planar_dimensions"], ft["cyl_centers"],
                   ft["cyl_normals"], ft["cyl_axes"], ft["cyl_radii"], ft["cyl_heights"], ft["cyl_opening_angles"])


def aim_points(targets: Targets, tidx: torch.Tensor) -> torch.Tensor:
    """``SolarTower.get_centers_of_target_areas`` (``artist/field/solar_tower.py:160-188``): plane centre, or
    cylinder centre + radius * normal; w = 1."""
    out = torch.zeros(tidx.shape[0], 4)
    for i, t in enumerate(tidx.tolist()):
        if t < targets.n_planar:
            out[i] = targets.planar_centers[t]
        else:
            k = t - targets.n_planar
            out[i] = targets.cyl_centers[k] + targets.cyl_radii[k] * targets.cyl_normals[k]
    out[:, 3] = 1.0
    return out


# --------------------------------------------------------------------------------------
# a16  sampler  (artist/raytracing/sampling.py:88-157)
# --------------------------------------------------------------------------------------


def sampler_indices(number_of_samples: int, number_of_active_heliostats: int,
                    world_size: int = 1, rank: int = 0) -> list[int]:
    """Heliostat ``h`` (all its replicas, contiguous) goes to rank ``h % active_ranks``."""
    active_ranks = min(number_of_active_heliostats, world_size)
    if rank >= active_ranks:
        return []
    per = number_of_samples // number_of_active_heliostats
    out: list[int] = []
    for h in range(number_of_active_heliostats):
        if h % active_ranks == rank:
            out.extend(range(h * per, h * per + per))
    return out


# --------------------------------------------------------------------------------------
# a8  sun distortions  (artist/scene/sun.py:96-119,224-234)
# --------------------------------------------------------------------------------------


def sun_distortions(number_of_rays: int, number_of_points: int, number_of_active_heliostats: int,
                    random_seed: int = 7, mean: float = 0.0, covariance: float = 4.3681e-06,
                    device: str | torch.device = "cpu") -> tuple[torch.Tensor, torch.Tensor]:
    """Reseed the *global* RNG and draw ``[N,R,P,2]`` from N(mean, cov*I); return (u, e) views."""
    m = torch.tensor([mean, mean], dtype=torch.float, device=device)
    c = torch.tensor([[covariance, 0], [0, covariance]], dtype=torch.float, device=device)
    dist = torch.distributions.MultivariateNormal(m, c)
    torch.manual_seed(random_seed)
    torch.cuda.manual_seed(random_seed)
    s = dist.sample((number_of_active_heliostats, number_of_rays, number_of_points))
    du, de = s.permute(3, 0, 1, 2)
    return du, de


# --------------------------------------------------------------------------------------
# a7  reflect  (artist/raytracing/geometry.py:32-41)
# --------------------------------------------------------------------------------------


def reflect(incident: torch.Tensor, normals: torch.Tensor) -> torch.Tensor:
    """``r = i - 2 (i.n) n`` with the 4-component dot (w components are 0)."""
    return incident - 2 * torch.sum(incident * normals, dim=-1, keepdim=True) * normals


# --------------------------------------------------------------------------------------
# a9  scatter  (artist/geometry/transforms.py:52-83, heliostat_ray_tracer.py:543-552)
# --------------------------------------------------------------------------------------


# cos / sin of the scatter angles.  The reference calls torch.cos / torch.sin (transforms.py:52-55); torch's CPU kernels
# (SLEEF, <= 1 ulp) return a value that is NOT the correctly rounded one for ~8.6 % of sun-shape angles.  Tests swap in
# ``correctly_rounded_trig`` to attribute differences between the reference and a faithful device implementation to
# exactly that (tests/test_gpu_large_mode_parity.py, bench.py's parity leg).
_scatter_cos, _scatter_sin = torch.cos, torch.sin


class correctly_rounded_trig:
    """Context manager: the scatter matrices use float64 cos / sin rounded once to float32."""

    def __enter__(self):
        global _scatter_cos, _scatter_sin
        self._saved = (_scatter_cos, _scatter_sin)
        _scatter_cos = lambda x: torch.cos(x.double()).to(x.dtype)
        _scatter_sin = lambda x: torch.sin(x.double()).to(x.dtype)
        return self

    def __exit__(self, *exc):
        global _scatter_cos, _scatter_sin
        _scatter_cos, _scatter_sin = self._saved
        return False


def rotate_distortions(e: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """Per-ray 4x4 rotation (first around up by ``u``, then around east by ``e``)."""
    if e.shape != u.shape:
        raise ValueError("e and u must have the same shape")
    ce, se, cu, su = _scatter_cos(e), _scatter_sin(e), _scatter_cos(u), _scatter_sin(u)
    m = torch.zeros(*e.shape, 4, 4, device=e.device)
    m[..., 0, 0] = cu
    m[..., 0, 1] = -su
    m[..., 1, 0] = ce * su
    m[..., 1, 1] = ce * cu
    m[..., 1, 2] = -se
    m[..., 2, 0] = se * su
    m[..., 2, 1] = se * cu
    m[..., 2, 2] = ce
    m[..., 3, 3] = 1.0
    return m


def scatter_rays(dist_u: torch.Tensor, dist_e: torch.Tensor, directions: torch.Tensor) -> torch.Tensor:
    """``[B,R,P]`` distortions x ``[B,P,4]`` preferred directions -> ``[B,R,P,4]``."""
    m = rotate_distortions(e=dist_e, u=dist_u)
    return (m @ directions.unsqueeze(1).unsqueeze(-1)).squeeze(-1)


# --------------------------------------------------------------------------------------
# a10  line-plane  (artist/raytracing/geometry.py:100-204)
# --------------------------------------------------------------------------------------


def line_plane_intersections(dirs, magnitudes, origins, centers, normals, dimensions, tidx, resolution):
    """Planar target hit -> (be, bu, t, intensity), each ``[B,R,P]``.

    ``resolution`` is the ``[E, U]`` integer tensor; note the hit uses *world* E and U
    components (vertical targets facing +-N) and flips E at the end.
    """
    resolution = resolution.to(dirs.device)
    tidx = tidx.long()
    d3 = dirs[..., :3]
    o3 = origins[..., :3]
    n3 = normals[tidx][..., :3]
    c3 = centers[tidx][..., :3]
    a = (d3 * n3[:, None, None, :]).sum(dim=-1)
    front = a < 0.0
    num = ((c3[:, None, :] - o3) * n3[:, None, :]).sum(dim=-1)[:, None, :]
    t = (num / torch.where(front, a, 1.0)) * front
    hit = o3[:, None, :, :] + d3 * t[..., None]
    inten = magnitudes * -a
    dims = dimensions[tidx]
    c4 = centers[tidx]
    te = hit[..., 0] + (dims[:, 0] / 2)[:, None, None] - c4[:, 0][:, None, None]
    tu = hit[..., 2] + (dims[:, 1] / 2)[:, None, None] - c4[:, 2][:, None, None]
    be = te / dims[:, 0, None, None] * (resolution[0] - 1)
    bu = tu / dims[:, 1, None, None] * (resolution[1] - 1)
    valid = (0 <= be) & (be <= resolution[0] - 1) & (0 <= bu) & (bu <= resolution[1] - 1) & front
    be = be * valid
    bu = bu * valid
    t = t * valid
    inten = inten * valid
    be = (resolution[0] - 1) - be
    return be, bu, t, inten


# --------------------------------------------------------------------------------------
# a11  line-cylinder  (artist/raytracing/geometry.py:287-445)
# --------------------------------------------------------------------------------------


def line_cylinder_intersections(dirs, magnitudes, origins, centers, normals, axes, radii, heights,
                                opening_angles, tidx, resolution):
    """Cylindrical sector hit -> (be, bu, t, intensity); no E flip; two data-dependent early exits."""
    resolution = resolution.to(dirs.device)
    tidx = tidx.long()
    o = origins[:, :, :3]
    d = dirs[:, :, :, :3]
    ax = axes[tidx][:, :3]
    nn = normals[tidx][:, :3]
    cc = centers[tidx][:, :3]
    rad = radii[tidx]
    hgt = heights[tidx]
    opn = opening_angles[tidx]
    uu = torch.cross(nn, ax, dim=-1)
    rot = torch.stack([uu, nn, ax], dim=1)
    ol = ((o - cc[:, None, :]) @ rot.transpose(1, 2))[:, None, :, :]
    dl = d @ rot.transpose(1, 2)[:, None, :, :]
    ox, oy = ol[..., 0], ol[..., 1]
    dx, dy = dl[..., 0], dl[..., 1]
    a = dx**2 + dy**2
    b = 2 * (ox * dx + oy * dy)
    c = (ox**2 + oy**2 - rad.view(-1, 1, 1) ** 2).repeat(1, dx.shape[1], 1)
    disc = b**2 - 4 * a * c
    hits = (disc >= 0) & (torch.abs(a) > 1e-8)
    zero = torch.zeros(d.shape[:3], device=d.device)
    if not torch.any(hits):
        return zero, zero, zero, zero
    sq = torch.sqrt(disc * hits + 1e-12)
    cand = torch.zeros(*d.shape[:3], 2, device=d.device)
    cand[..., 0] = (-b - sq) / (2 * a)
    cand[..., 1] = (-b + sq) / (2 * a)
    cand = torch.where(cand > 0, cand, torch.full_like(cand, torch.inf))
    t, _ = torch.min(cand, dim=-1)
    vd = torch.isfinite(t) & hits
    t = torch.where(vd, t, torch.zeros_like(t))
    if (t == 0.0).all():
        return zero, zero, zero, zero
    p = ol + t[..., None] * dl
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    nl = torch.stack([x, y, torch.zeros_like(x)], dim=-1)
    nl = nl / torch.norm(nl, dim=-1, keepdim=True)
    lam = (-dl * nl).sum(dim=-1).clamp(min=0.0)
    z = z + hgt.view(-1, 1, 1) / 2
    ang = torch.atan2(y, x) - (torch.atan2(nn[:, 1].view(-1, 1, 1), nn[:, 0].view(-1, 1, 1))
                               - (opn.view(-1, 1, 1) / 2))
    on = (z >= 0) & (z <= hgt.view(-1, 1, 1)) & (ang >= 0) & (ang <= opn.view(-1, 1, 1))
    bu = z / hgt.view(-1, 1, 1) * (resolution[1] - 1)
    be = ang / opn.view(-1, 1, 1) * (resolution[0] - 1)
    be = be * on * vd
    bu = bu * on * vd
    t = t * on * vd
    inten = magnitudes * lam * on * vd
    return be, bu, t, inten


# --------------------------------------------------------------------------------------
# a13  bilinear splat  (artist/raytracing/heliostat_ray_tracer.py:648-778)
# --------------------------------------------------------------------------------------


def bilinear_splatting(be: torch.Tensor, bu: torch.Tensor, inten: torch.Tensor, resolution) -> torch.Tensor:
    """4-tap deposit in tap order 1,2,3,4 over rays in flat (r,p) order; rows flipped at the end."""
    width = int(resolution[0])
    height = int(resolution[1])
    n = inten.shape[0]
    be = be.reshape(n, -1)
    bu = bu.reshape(n, -1)
    inten = inten.reshape(n, -1)
    ie = be.long()
    iu = bu.long()
    wle = ie + 1 - be
    wlu = iu + 1 - bu
    whe = be - ie
    whu = bu - iu
    p1 = wle * whu * inten
    p2 = whe * whu * inten
    p3 = whe * wlu * inten
    p4 = wle * wlu * inten
    on = (0 <= ie) & (ie + 1 < width) & (0 <= iu) & (iu + 1 < height)
    flat = torch.zeros((n, height * width), device=inten.device)
    i1 = (iu + 1) * width + ie
    i2 = (iu + 1) * width + ie + 1
    i3 = iu * width + ie + 1
    i4 = iu * width + ie
    for idx in (i1, i2, i3, i4):
        idx[~on] = 0
    flat.scatter_add_(1, i1, p1 * on)
    flat.scatter_add_(1, i2, p2 * on)
    flat.scatter_add_(1, i3, p3 * on)
    flat.scatter_add_(1, i4, p4 * on)
    return torch.flip(flat.view(n, height, width), [1])


# --------------------------------------------------------------------------------------
# a12  blocking  (artist/raytracing/blocking.py:123-209, 212-354, 832-995)
# --------------------------------------------------------------------------------------


def blocking_primitives(aligned_surface_points: torch.Tensor):
    """Rectangle per heliostat from 4 fixed corner point indices (assumes 4 square facets, 2x2).

    ``blocking.py:176-209``: corners 0..3 = points[P/2], [sqrt(P/4)-1], [P/2-1], [P-sqrt(P/4)];
    spans u = c1-c0, v = c3-c0; normal = normalize(u x v).
    """
    p = aligned_surface_points.shape[1]
    q = int(math.sqrt(p / 4))
    idx = torch.tensor([p // 2, q - 1, p // 2 - 1, p - q], device=aligned_surface_points.device)
    corners = aligned_surface_points[:, idx, :3]
    span_u = corners[:, 1] - corners[:, 0]
    span_v = corners[:, 3] - corners[:, 0]
    normals = torch.nn.functional.normalize(torch.linalg.cross(span_u, span_v), dim=-1)
    return corners, torch.stack([span_u, span_v], dim=1), normals


def soft_ray_blocking_mask(origins, dirs, corners, spans, normals, epsilon: float = 1e-12, softness: float = 1000.0,
                           alpha: float = 100.0, ray_origin_offset: float = 0.05):
    """Soft differentiable ray/rectangle blocking with Beer-Lambert accumulation (``blocking.py:289-354``).

    ``origins [B,P,>=3]``, ``dirs [B,R,P,>=3]``, primitives ``corners [K,4,>=3]``, ``spans [K,2,>=3]``,
    ``normals [K,>=3]`` -> ``blocked [B,R,P]`` in [0, 1).
    """
    o = origins[:, None, :, None, :3]
    d = dirs[:, :, :, None, :3]
    c0 = corners[None, None, None, :, 0, :3]
    su = spans[None, None, None, :, 0, :3]
    sv = spans[None, None, None, :, 1, :3]
    nn = normals[None, None, None, :, :3]
    den = torch.sum(d * nn, dim=-1)
    den = torch.where(den.abs() < epsilon, torch.where(den >= 0, epsilon, -epsilon), den)
    t = torch.sum((c0 - o) * nn, dim=-1) / den
    front = torch.sigmoid(softness * (t - ray_origin_offset))
    off = (o + t[..., None] * d) - c0
    uu, vv, uv = torch.sum(su * su, dim=-1), torch.sum(sv * sv, dim=-1), torch.sum(su * sv, dim=-1)
    pu, pv = torch.sum(off * su, dim=-1), torch.sum(off * sv, dim=-1)
    det = uu * vv - uv * uv
    det = torch.where(det.abs() < epsilon, torch.sign(det) * epsilon, det)
    u = (pu * vv - pv * uv) / det
    v = (pv * uu - pu * uv) / det
    inside = (torch.sigmoid(softness * u) * torch.sigmoid(softness * (1 - u))
              * torch.sigmoid(softness * v) * torch.sigmoid(softness * (1 - v)))
    sigma = (inside * front).clamp(0.0, 1.0)
    return 1.0 - torch.exp(-(alpha * torch.sum(sigma, dim=-1)))


@torch.no_grad()
def blocking_filter(origins, dirs, corners, ray_owner, t_target):
    """The SET of primitives whose axis-aligned box is hit by any ray of the batch before the ray reaches its
    target, self-hits removed - what the reference's LBVH traversal returns (``blocking.py:832-995``); brute force
    over primitives here.  ``origins [B,P,>=3]``, ``dirs [B,R,P,>=3]``, ``ray_owner [B]``, ``t_target [B,R,P]``."""
    b, r, p = dirs.shape[:3]
    o = origins[:, None, :, :3].expand(-1, r, -1, -1).reshape(-1, 3)
    inv = 1.0 / (dirs[..., :3].reshape(-1, 3) + 1e-12)
    tt = t_target.reshape(-1)
    owner = ray_owner.repeat_interleave(r * p)
    c3 = corners[..., :3]
    keep = []
    for k in range(c3.shape[0]):
        lo, hi = c3[k].min(dim=0).values, c3[k].max(dim=0).values
        d0, d1 = (lo - o) * inv, (hi - o) * inv
        entry = torch.minimum(d0, d1).amax(dim=-1)
        exit_ = torch.maximum(d0, d1).amin(dim=-1)
        hit = (exit_ >= entry) & (exit_ > 1e-6) & (entry <= tt) & (owner != k)
        if hit.any():
            keep.append(k)
    return torch.tensor(keep, dtype=torch.long)


# --------------------------------------------------------------------------------------
# a7-a15  trace_rays  (artist/raytracing/heliostat_ray_tracer.py:220-508, 563-608)
# --------------------------------------------------------------------------------------


def trace_rays(points, normals, incident, dist_u, dist_e, target_idx, targets: Targets,
               resolution=(256, 256), ray_magnitude: float = 1.0, ray_extinction_factor: float = 0.0,
               mirror_reflectivity: float = 0.935, sample_indices: list[int] | None = None,
               batch_size: int = 100, blocking: dict | None = None):
    """Full planar/cylindrical trace; ``blocking = dict(corners[H,4,3+], spans[H,2,3+], normals[H,3+],
    sample_to_blocker[N])`` switches on the blocking branch of ``heliostat_ray_tracer.py:445-480`` (per-batch
    primitive filter + soft mask), ``None`` = ``blocking_active=False``.

    Returns ``(flux[N,U,E], intercept[N], on_target[N], blocking[N])`` like the reference
    (``:508``) except that rows of samples not in ``sample_indices`` are ZERO instead of
    uninitialised (consciously fixed).
    """
    res = torch.as_tensor(resolution)
    n = points.shape[0]
    r = dist_u.shape[1]
    p = points.shape[1]
    width, height = int(res[0]), int(res[1])
    if sample_indices is None:
        sample_indices = list(range(n))
    flux = torch.zeros(n, height, width, device=points.device)
    intercept = torch.zeros(n, device=points.device)
    on_target = torch.zeros(n, device=points.device)
    blocking_f = torch.zeros(n, device=points.device)
    ref = reflect(incident.unsqueeze(1), normals)
    tidx_all = target_idx.long()
    for s in range(0, len(sample_indices), batch_size):
        b = torch.as_tensor(sample_indices[s:s + batch_size], dtype=torch.long, device=points.device)
        dirs = scatter_rays(dist_u[b], dist_e[b], ref[b])
        mags = torch.full(dirs.shape[:3], ray_magnitude, device=points.device)
        tb = tidx_all[b]
        planar = tb < targets.n_planar
        be = torch.zeros(len(b), r, p, device=points.device)
        bu = torch.zeros_like(be)
        tt = torch.zeros_like(be)
        lam = torch.zeros_like(be)
        if planar.sum() > 0:
            be[planar], bu[planar], tt[planar], lam[planar] = line_plane_intersections(
                dirs[planar], mags[planar], points[b][planar], targets.planar_centers,
                targets.planar_normals, targets.planar_dimensions, tb[planar], res)
        if (~planar).sum() > 0:
            be[~planar], bu[~planar], tt[~planar], lam[~planar] = line_cylinder_intersections(
                dirs[~planar], mags[~planar], points[b][~planar], targets.cyl_centers, targets.cyl_normals,
                targets.cyl_axes, targets.cyl_radii, targets.cyl_heights, targets.cyl_opening_angles,
                tb[~planar] - targets.n_planar, res)
        blocked = torch.zeros_like(be)
        if blocking is not None:
            keep = blocking_filter(points[b], dirs, blocking["corners"], blocking["sample_to_blocker"][b], tt)
            if keep.numel() > 0:
                blocked = soft_ray_blocking_mask(points[b], dirs, blocking["corners"][keep], blocking["spans"][keep],
                                                 blocking["normals"][keep])
        inten = lam * (1 - blocked) * (1 - ray_extinction_factor) * mirror_reflectivity
        flux[b] = bilinear_splatting(be, bu, inten, res)
        rp = r * p
        on_target[b] = (lam > 0).sum((1, 2)) / rp
        blocking_f[b] = (blocked < 1e-3).sum((1, 2)) / rp
        intercept[b] = (inten > 0).sum((1, 2)) / rp
    return flux, intercept, on_target, blocking_f


def ray_pixel_coordinates(points, normals, incident, dist_u, dist_e, target_idx, targets: Targets,
                          resolution=(256, 256), ray_magnitude: float = 1.0):
    """Per-ray (be, bu, t, lambert) ``[N,R,P]`` - the parity-critical intermediates."""
    res = torch.as_tensor(resolution)
    ref = reflect(incident.unsqueeze(1), normals)
    dirs = scatter_rays(dist_u, dist_e, ref)
    mags = torch.full(dirs.shape[:3], ray_magnitude)
    tb = target_idx.long()
    planar = tb < targets.n_planar
    n, r, p = dist_u.shape
    out = [torch.zeros(n, r, p) for _ in range(4)]
    if planar.any():
        vals = line_plane_intersections(dirs[planar], mags[planar], points[planar], targets.planar_centers,
                                        targets.planar_normals, targets.planar_dimensions, tb[planar], res)
        for o, v in zip(out, vals):
            o[planar] = v
    if (~planar).any():
        vals = line_cylinder_intersections(dirs[~planar], mags[~planar], points[~planar], targets.cyl_centers,
                                           targets.cyl_normals, targets.cyl_axes, targets.cyl_radii,
                                           targets.cyl_heights, targets.cyl_opening_angles,
                                           tb[~planar] - targets.n_planar, res)
        for o, v in zip(out, vals):
            o[~planar] = v
    return tuple(out)


def bitmaps_per_target(bitmaps: torch.Tensor, target_idx: torch.Tensor, n_targets: int) -> torch.Tensor:
    """``heliostat_ray_tracer.py:593-608``: per target, sum of the bitmaps aimed at it."""
    out = torch.zeros(n_targets, bitmaps.shape[1], bitmaps.shape[2], device=bitmaps.device)
    for t in range(n_targets):
        m = target_idx == t
        if m.any():
            out[t] = bitmaps[m].sum(dim=0)
    return out


# --------------------------------------------------------------------------------------
# a1  NURBS  (artist/nurbs/surfaces.py:98-155,198-207,325-417,592-687; nurbs/utils.py:7-49)
# --------------------------------------------------------------------------------------


def nurbs_evaluation_grid(pu: int, pv: int, epsilon: float = 1e-7) -> torch.Tensor:
    return torch.cartesian_prod(torch.linspace(epsilon, 1 - epsilon, pu), torch.linspace(epsilon, 1 - epsilon, pv))


def planar_control_points(cu: int, cv: int, canting: torch.Tensor) -> torch.Tensor:
    """``nurbs/utils.py:52-121``: flat control net spanning +-|canting vector| per facet; ``[F,cu,cv,3]``."""
    f = canting.shape[0]
    cp = torch.zeros(f, cu, cv, 3, dtype=canting.dtype)
    dims = torch.norm(canting, dim=2)
    ul = torch.linspace(0, 1, cu, dtype=canting.dtype)
    vl = torch.linspace(0, 1, cv, dtype=canting.dtype)
    uc = -dims[:, 0, None] + 2 * dims[:, 0, None] * ul
    vc = -dims[:, 1, None] + 2 * dims[:, 1, None] * vl
    cp[..., 0] = uc[:, :, None]
    cp[..., 1] = vc[:, None, :]
    return cp


def uniform_knots(n_ctrl: int, degree: int) -> torch.Tensor:
    """Clamped uniform knot vector (``surfaces.py:128-147``)."""
    k = torch.zeros(n_ctrl + degree + 1)
    k[degree:-degree] = torch.linspace(0, 1, n_ctrl - degree + 1)
    k[-degree:] = 1
    return k


def _basis(x: torch.Tensor, knots: torch.Tensor, span: torch.Tensor, degree: int):
    """Basis values and first derivatives (NURBS Book A2.3 restricted to k<=1; ``surfaces.py:325-415``)."""
    ndu = [[None] * (degree + 1) for _ in range(degree + 1)]
    ndu[0][0] = torch.ones_like(x)
    left = [None] * (degree + 1)
    right = [None] * (degree + 1)
    for j in range(1, degree + 1):
        left[j] = x - knots[span - j + 1]
        right[j] = knots[span + j] - x
        saved = torch.zeros_like(x)
        for r in range(j):
            ndu[j][r] = right[r + 1] + left[j - r]
            tmp = ndu[r][j - 1] / ndu[j][r]
            ndu[r][j] = saved + right[r + 1] * tmp
            saved = left[j - r] * tmp
        ndu[j][j] = saved
    n0 = [ndu[j][degree] for j in range(degree + 1)]
    n1 = []
    pk = degree - 1
    for r in range(degree + 1):
        d = torch.zeros_like(x)
        a10 = None
        if r >= 1:
            a10 = 1.0 / ndu[pk + 1][r - 1]
            d = a10 * ndu[r - 1][pk]
        if r <= pk:
            a11 = -1.0 / ndu[pk + 1][r]
            d = d + a11 * ndu[r][pk]
        n1.append(d * degree)
    return n0, n1


def perform_canting(canting: torch.Tensor, data: torch.Tensor) -> torch.Tensor:
    """``geometry/transforms.py:321-347``: orthonormal basis from the two canting vectors, ``data @ R^T``."""
    e = torch.nn.functional.normalize(canting[:, :, 0, :3], dim=-1)
    n = canting[:, :, 1, :3]
    u = torch.nn.functional.normalize(torch.linalg.cross(e, n, dim=-1), dim=-1, eps=1e-8)
    no = torch.nn.functional.normalize(torch.linalg.cross(u, e, dim=-1), dim=-1, eps=1e-8)
    rot = torch.zeros(*canting.shape[:2], 4, 4, dtype=data.dtype, device=data.device)
    rot[:, :, :3, 0] = e
    rot[:, :, :3, 1] = no
    rot[:, :, :3, 2] = u
    rot[:, :, 3, 3] = 1.0
    return data @ rot.mT


def nurbs_points_and_normals(control_points: torch.Tensor, degree_u: int, degree_v: int,
                             evaluation_points: torch.Tensor, canting: torch.Tensor | None = None,
                             facet_translations: torch.Tensor | None = None):
    """``control_points [N,F,cu,cv,3]``, ``evaluation_points [N,F,Pf,2]`` -> points, normals ``[N,F,Pf,4]``."""
    n, f, cu, cv, _ = control_points.shape
    ku = uniform_knots(cu, degree_u).to(control_points.device)
    kv = uniform_knots(cv, degree_v).to(control_points.device)
    xu = evaluation_points[..., 0]
    xv = evaluation_points[..., 1]
    su = torch.floor(xu * (cu - degree_u)).long() + degree_u
    sv = torch.floor(xv * (cv - degree_v)).long() + degree_v
    nu0, nu1 = _basis(xu, ku, su, degree_u)
    nv0, nv1 = _basis(xv, kv, sv, degree_v)
    cp = torch.cat([control_points, torch.ones(n, f, cu, cv, 1, device=control_points.device)], dim=-1)
    bi = torch.arange(n, device=cp.device).view(n, 1, 1).expand(n, f, xu.shape[2])
    fi = torch.arange(f, device=cp.device).view(1, f, 1).expand(n, f, xu.shape[2])

    def temps(bu_list):
        # temp[s] = sum_r N_u[r] * P[span_u - p + r, span_v - q + s]   (surfaces.py:592-605)
        out = []
        for s in range(degree_v + 1):
            tmp = torch.zeros(n, f, xu.shape[2], 4, device=cp.device)
            for r in range(degree_u + 1):
                tmp = tmp + bu_list[r].unsqueeze(-1) * cp[bi, fi, su - degree_u + r, sv - degree_v + s]
            out.append(tmp)
        return out

    def combine(tmp, bv_list):
        out = torch.zeros(n, f, xu.shape[2], 4, device=cp.device)
        for s in range(degree_v + 1):
            out = out + bv_list[s].unsqueeze(-1) * tmp[s]
        return out

    t_k0 = temps(nu0)
    s00 = combine(t_k0, nv0)
    s01 = combine(t_k0, nv1)
    s10 = combine(temps(nu1), nv0)
    normals = torch.nn.functional.normalize(torch.linalg.cross(s10[..., :3], s01[..., :3]), dim=3)
    points = torch.cat([s00[..., :3] / s00[..., 3:4], torch.ones(n, f, xu.shape[2], 1, device=cp.device)], dim=3)
    normals = torch.cat([normals, torch.zeros(n, f, xu.shape[2], 1, device=cp.device)], dim=3)
    if canting is not None:
        points = perform_canting(canting, points) + facet_translations.reshape(n, f, 1, 4)
        normals = perform_canting(canting, normals)
    return points, normals


# --------------------------------------------------------------------------------------
# a3-a5  kinematics + actuators
#   (artist/field/kinematics_rigid_body.py:194-324,326-508,540-634; actuators_linear.py:79-370)
# --------------------------------------------------------------------------------------


def _rot_e(a):
    m = torch.zeros(a.shape[0], 4, 4, device=a.device)
    c, s = torch.cos(a), torch.sin(a)
    m[:, 0, 0] = 1
    m[:, 1, 1] = c
    m[:, 1, 2] = -s
    m[:, 2, 1] = s
    m[:, 2, 2] = c
    m[:, 3, 3] = 1
    return m


def _rot_n(a):
    m = torch.zeros(a.shape[0], 4, 4, device=a.device)
    c, s = torch.cos(a), torch.sin(a)
    m[:, 0, 0] = c
    m[:, 0, 2] = -s
    m[:, 1, 1] = 1
    m[:, 2, 0] = s
    m[:, 2, 2] = c
    m[:, 3, 3] = 1
    return m


def _rot_u(a):
    m = torch.zeros(a.shape[0], 4, 4, device=a.device)
    c, s = torch.cos(a), torch.sin(a)
    m[:, 0, 0] = c
    m[:, 0, 1] = -s
    m[:, 1, 0] = s
    m[:, 1, 1] = c
    m[:, 2, 2] = 1
    m[:, 3, 3] = 1
    return m


def _trans(e, n, u):
    m = torch.zeros(e.shape[0], 4, 4, device=e.device)
    m[:, 0, 0] = 1
    m[:, 1, 1] = 1
    m[:, 2, 2] = 1
    m[:, 3, 3] = 1
    m[:, 0, 3] = e
    m[:, 1, 3] = n
    m[:, 2, 3] = u
    return m


def initial_orientation_offset() -> torch.Tensor:
    """Rotation up(0,0,1) -> south(0,-1,0): ``rotate_e(pi/2)`` for these two fixed vectors
    (``kinematics_rigid_body.py:178-190`` with ``rotations.decompose_rotations``)."""
    a = torch.nn.functional.normalize(torch.tensor([[0.0, 0.0, 1.0]]))
    b = torch.nn.functional.normalize(torch.tensor([0.0, -1.0, 0.0]), dim=0).unsqueeze(0)
    r = torch.nn.functional.normalize(torch.linalg.cross(a, b))
    theta = torch.arccos(torch.clamp(a @ b.T, -1.0, 1.0))
    comp = theta * r
    return _rot_e(comp[:, 0]) @ _rot_n(comp[:, 1]) @ _rot_u(comp[:, 2])


@dataclass
class Kin:
    """Active per-heliostat kinematic tensors (``kinematics_rigid_body.py:133-172``)."""

    positions: torch.Tensor            # [N,4]
    translation_deviations: torch.Tensor  # [N,9]
    rotation_deviations: torch.Tensor     # [N,4]
    actuator_non_optimizable: torch.Tensor  # [N,7,2]
    actuator_optimizable: torch.Tensor      # [N,2,2] (linear) or empty (ideal)
    linear: bool = True


def _softplus100(x):
    return torch.nn.functional.softplus(x, beta=100)


def _linear_params(k: Kin):
    eps = 1e-6
    inc = _softplus100(k.actuator_non_optimizable[:, 4]) + eps
    off = _softplus100(k.actuator_non_optimizable[:, 5]) + eps
    rad = _softplus100(k.actuator_non_optimizable[:, 6]) + eps
    init_angle = k.actuator_optimizable[:, 0]
    init_stroke = _softplus100(k.actuator_optimizable[:, 1]) + eps
    return inc, off, rad, init_angle, init_stroke


def _linear_abs_angles(k: Kin, motor):
    eps = 1e-6
    inc, off, rad, _, s0 = _linear_params(k)
    stroke = motor / inc + s0
    stroke = torch.clamp(stroke, min=(off - rad).abs() + eps, max=off + rad - eps)
    div = (off**2 + rad**2 - stroke**2) / (2.0 * off * rad)
    return torch.arccos(torch.clamp(div, min=-1.0 + 1e-6, max=1.0 - 1e-6))


def motor_positions_to_angles(k: Kin, motor: torch.Tensor) -> torch.Tensor:
    if not k.linear:
        return motor
    _, _, _, a0, _ = _linear_params(k)
    ab = _linear_abs_angles(k, motor)
    ab0 = _linear_abs_angles(k, torch.zeros_like(motor))
    delta = ab0 - ab
    cw = k.actuator_non_optimizable[:, 1]
    return a0 + delta * (cw == 1) - delta * (cw == 0)


def angles_to_motor_positions(k: Kin, angles: torch.Tensor) -> torch.Tensor:
    if not k.linear:
        return angles
    eps = 1e-6
    inc, off, rad, a0, s0 = _linear_params(k)
    cw = k.actuator_non_optimizable[:, 1]
    delta = torch.where(cw == 1, angles - a0, a0 - angles)
    ab0 = _linear_abs_angles(k, torch.zeros_like(angles))
    ia = ab0 - delta
    cosv = torch.clamp(torch.cos(ia), -1.0 + 1e-6, 1.0 - 1e-6)
    stroke = torch.sqrt(off**2 + rad**2 - 2.0 * off * rad * cosv)
    stroke = torch.clamp(stroke, min=(off - rad).abs() + eps, max=off + rad - eps)
    return (stroke - s0) * inc


def orientations_from_motor_positions_raw(k: Kin, motor: torch.Tensor) -> torch.Tensor:
    """``_compute_orientations_from_motor_positions`` (``:194-324``) - WITHOUT the final offset."""
    ang = motor_positions_to_angles(k, motor)
    td, rd = k.translation_deviations, k.rotation_deviations
    o = torch.eye(4, device=motor.device)[None] @ _trans(k.positions[:, 0], k.positions[:, 1], k.positions[:, 2])
    j1 = _rot_n(rd[:, 0]) @ _rot_u(rd[:, 1]) @ _trans(td[:, 0], td[:, 1], td[:, 2]) @ _rot_e(ang[:, 0])
    j2 = _rot_e(rd[:, 2]) @ _rot_n(rd[:, 3]) @ _trans(td[:, 3], td[:, 4], td[:, 5]) @ _rot_u(ang[:, 1])
    jr = torch.zeros(motor.shape[0], 2, 4, 4, device=motor.device)
    jr[:, 0] = j1
    jr[:, 1] = j2
    return o @ jr[:, 0] @ jr[:, 1] @ _trans(td[:, 6], td[:, 7], td[:, 8])


def motor_positions_to_orientations(k: Kin, motor: torch.Tensor) -> torch.Tensor:
    """``:510-538``."""
    return orientations_from_motor_positions_raw(k, motor) @ initial_orientation_offset().to(motor.device)


def motor_positions_from_normal(k: Kin, normals: torch.Tensor, epsilon: float = 1e-8) -> torch.Tensor:
    """Closed-form two-solution inverse kinematics (``:373-508``)."""
    rd = k.rotation_deviations
    f1 = _rot_n(rd[:, 0]) @ _rot_u(rd[:, 1])
    f2 = _rot_e(rd[:, 2]) @ _rot_n(rd[:, 3])
    npr = (f1.transpose(-1, -2) @ normals[:, :, None])[:, :, 0]
    f00, f01 = f2[:, 0, 0], f2[:, 0, 1]
    den = torch.sqrt(f00**2 + f01**2)
    phi = torch.atan2(-f01, f00)
    ratio = torch.clamp(npr[:, 0] / (den + epsilon), -1.0 + epsilon, 1.0 - epsilon)
    s1 = torch.arcsin(ratio) - phi
    s2 = torch.pi - torch.arcsin(ratio) - phi
    s1 = torch.atan2(torch.sin(s1), torch.cos(s1))
    s2 = torch.atan2(torch.sin(s2), torch.cos(s2))
    south = torch.tensor([0.0, -1.0, 0.0, 0.0], device=normals.device)

    def first(sa):
        v = f2 @ _rot_u(sa) @ south
        a = torch.atan2(v[:, 1] * npr[:, 2] - v[:, 2] * npr[:, 1], v[:, 1] * npr[:, 1] + v[:, 2] * npr[:, 2])
        return torch.atan2(torch.sin(a), torch.cos(a))

    m1 = angles_to_motor_positions(k, torch.stack([first(s1), s1], dim=-1))
    m2 = angles_to_motor_positions(k, torch.stack([first(s2), s2], dim=-1))
    lo, hi = k.actuator_non_optimizable[:, 2], k.actuator_non_optimizable[:, 3]
    ok1 = ((m1 >= lo) & (m1 <= hi)).all(dim=1)
    return torch.where(ok1[:, None], m1, m2)


def incident_ray_directions_to_orientations(k: Kin, incident: torch.Tensor, aim_points: torch.Tensor,
                                            max_num_iterations: int = 4, min_eps: float = 0.0001):
    """<=4 fixed-point iterations, stop when ALL heliostats converged (``:576-634``).

    Returns ``(orientations[N,4,4], motor_positions[N,2])``.
    """
    n = incident.shape[0]
    motor = torch.zeros(n, 2, device=incident.device)
    south = torch.tensor([0.0, -1.0, 0.0, 0.0], device=incident.device)
    origin = torch.tensor([0.0, 0.0, 0.0, 1.0], device=incident.device)
    last = None
    for _ in range(max_num_iterations):
        o = orientations_from_motor_positions_raw(k, motor)
        cn = o @ south
        co = o @ origin
        want_ref = torch.nn.functional.normalize(aim_points[:, :3] - co[:, :3], p=2, dim=1, eps=1e-8)
        want_n = torch.nn.functional.normalize(-incident[:, :3] + want_ref, p=2, dim=1, eps=1e-8)
        want_n = torch.cat([want_n, torch.zeros(n, 1, device=incident.device)], dim=1)
        loss = torch.abs(want_n - cn).mean(dim=-1)
        if last is not None and torch.all(torch.abs(last - loss) <= min_eps):
            break
        last = loss
        motor = motor_positions_from_normal(k, want_n)
    return o @ initial_orientation_offset().to(incident.device), motor


def align_surfaces(points: torch.Tensor, normals: torch.Tensor, orientations: torch.Tensor):
    """``heliostat_group_rigid_body.py:217-222``: ``[N,P,4] @ O^T``."""
    return points @ orientations.transpose(1, 2), normals @ orientations.transpose(1, 2)


# --------------------------------------------------------------------------------------
# step after the path: flux-bitmap centre of mass and crop  (artist/flux/bitmap.py:12-55, 121-246)
# --------------------------------------------------------------------------------------


def flux_center_of_mass(bitmaps: torch.Tensor) -> torch.Tensor:
    """``[B,U,E]`` -> ``[B,2]`` (e, u) centre of mass in pixel units (``bitmap.py:37-55``)."""
    _, height, width = bitmaps.shape
    norm = bitmaps / (bitmaps.sum(dim=(1, 2), keepdim=True) + 1e-8)
    e = torch.linspace(0, width - 1, width)
    u = torch.linspace(0, height - 1, height)
    ug, eg = torch.meshgrid(u, e, indexing="ij")
    return torch.stack([(eg * norm).sum(dim=(1, 2)), (ug * norm).sum(dim=(1, 2))], dim=1)


def crop_flux_around_center(bitmaps: torch.Tensor, target_dimensions: torch.Tensor, crop_width: float = 6,
                            crop_height: float = 6) -> torch.Tensor:
    """Centre-of-mass crop (``bitmap.py:157-246``): ``target_dimensions [B,2]`` = (width, height) of each bitmap's
    target area in metres; affine_grid + grid_sample (bilinear, align_corners=True, zero padding)."""
    n, height, width = bitmaps.shape
    norm = bitmaps / (bitmaps.sum(dim=(2, 1), keepdim=True) + 1e-8)
    y = torch.linspace(-1, 1, height)
    x = torch.linspace(-1, 1, width)
    yg, xg = torch.meshgrid(y, x, indexing="ij")
    cx = (xg * norm).sum(dim=(2, 1))
    cy = (yg * norm).sum(dim=(2, 1))
    w = target_dimensions[:, 0].clamp(min=1e-8)
    h = target_dimensions[:, 1].clamp(min=1e-8)
    theta = torch.zeros(n, 2, 3)
    theta[:, 0, 0] = crop_width / w
    theta[:, 1, 1] = crop_height / h
    theta[:, 0, 2] = cx
    theta[:, 1, 2] = cy
    grid = torch.nn.functional.affine_grid(theta, size=[n, 1, height, width], align_corners=True)
    return torch.nn.functional.grid_sample(bitmaps[:, None], grid, align_corners=True, padding_mode="zeros")[:, 0]


# --------------------------------------------------------------------------------------------
# bitmap losses  (artist/optim/loss.py:251-319 PixelLoss, :322-410 KLDivergenceLoss)
# --------------------------------------------------------------------------------------------
def pixel_loss(prediction: torch.Tensor, ground_truth: torch.Tensor, reduction_dimensions=(1, 2)) -> torch.Tensor:
    """``loss.py:317-319``: summed squared error per sample / total ground-truth intensity of the sample."""
    return torch.nn.functional.mse_loss(prediction, ground_truth, reduction="none").sum(dim=reduction_dimensions) / \
        ground_truth.sum(dim=(1, 2))


def kl_divergence_loss(prediction: torch.Tensor, ground_truth: torch.Tensor, reduction_dimensions=(1, 2)) -> torch.Tensor:
    """``loss.py:389-410``: both bitmaps L1-normalised (eps), shifted by eps, ``KLDivLoss(log_target=True)`` of the logs."""
    eps = 1e-12
    p = torch.nn.functional.normalize(ground_truth, p=1, dim=(1, 2), eps=eps)
    q = torch.nn.functional.normalize(prediction, p=1, dim=(1, 2), eps=eps)
    loss = torch.nn.functional.kl_div(torch.log(q + eps), torch.log(p + eps), reduction="none", log_target=True)
    return loss.sum(dim=reduction_dimensions)
