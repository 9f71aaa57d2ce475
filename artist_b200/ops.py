"""Thin torch layer over the C ABI: tensors in, ``data_ptr()`` + current CUDA stream out, with
explicit ``autograd.Function`` backwards that call the hand-written backward kernels.

Nothing here computes on the CPU; every function raises if its tensors are not CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import ABI_VERSION


_warned_no_graph: set = set()


def warn_no_graph(what: str, hint: str = "") -> None:
    """One warning per call site: an input requires grad but this step returns tensors without an autograd graph."""
    if what in _warned_no_graph:
        return
    _warned_no_graph.add(what)
    import warnings

    warnings.warn(f"artist_b200: {what} is forward-only here (its inputs require grad, the result carries no autograd graph); "
                  f"gradients flow through HeliostatRayTracer.trace_rays / the fused ops instead. {hint}", stacklevel=3)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """Handle of torch's current CUDA stream (the raw getter is ~20x cheaper than building a ``torch.cuda.Stream`` object:
    at small sample counts the step is bound by host time, tools/host_profile.py)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.Ab200Error(f"{name} must be a CUDA tensor (artist_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.Ab200Error(f"{name} must be a CUDA tensor (artist_b200 has no CPU path)")
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def _p(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


@dataclass
class TargetTensors:
    """Device-resident target-area SoA (``artist/field/tower_target_areas_*.py``)."""

    planar_centers: torch.Tensor
    planar_normals: torch.Tensor
    planar_dims: torch.Tensor
    cyl_centers: torch.Tensor
    cyl_normals: torch.Tensor
    cyl_axes: torch.Tensor
    cyl_radii: torch.Tensor
    cyl_heights: torch.Tensor
    cyl_opening: torch.Tensor

    @property
    def n_planar(self) -> int:
        return int(self.planar_centers.shape[0])

    @property
    def n_cyl(self) -> int:
        return int(self.cyl_centers.shape[0])

    def struct(self) -> _lib.Targets:
        return _lib.Targets(self.n_planar, self.n_cyl, _p(self.planar_centers), _p(self.planar_normals),
                            _p(self.planar_dims), _p(self.cyl_centers), _p(self.cyl_normals), _p(self.cyl_axes),
                            _p(self.cyl_radii), _p(self.cyl_heights), _p(self.cyl_opening))

    @classmethod
    def from_solar_tower(cls, solar_tower, device) -> "TargetTensors":
        planar, cyl = solar_tower.target_areas[0], solar_tower.target_areas[1]

        def g(x, shape):
            x = torch.as_tensor(x, dtype=torch.float32, device=device).reshape(shape)
            return x.contiguous()

        return cls(g(planar.centers, (-1, 4)), g(planar.normals, (-1, 4)), g(planar.dimensions, (-1, 2)),
                   g(cyl.centers, (-1, 4)), g(cyl.normals, (-1, 4)), g(cyl.axes, (-1, 4)), g(cyl.radii, (-1,)),
                   g(cyl.heights, (-1,)), g(cyl.opening_angles, (-1,)))


@dataclass
class TraceOptions:
    res_e: int = 256
    res_u: int = 256
    ray_magnitude: float = 1.0
    ray_extinction_factor: float = 0.0
    mirror_reflectivity: float = 0.935
    scatter_sigma: float = 0.0
    trig_mode: int = _lib.TRIG_POLY
    fp32_accumulate: bool = False
    one_cta_per_sample: bool = False   # force the kernels a full field runs (AB200_FLAG_ONE_CTA_PER_SAMPLE)


@dataclass
class BlockingInputs:
    """Rectangle primitives of ALL heliostats (``artist/raytracing/blocking.py:123-209``) + what the candidate search
    needs per active sample.  ``corners/spans/normals`` may require grad (gradients flow to the blockers' geometry)."""

    corners: torch.Tensor            # [H,4,>=3]
    spans: torch.Tensor              # [H,2,>=3]
    normals: torch.Tensor            # [H,>=3]
    sample_to_blocker: torch.Tensor  # [N] int32: primitive row of each active sample's own heliostat
    aim_points: torch.Tensor         # [N,4]
    target_radius: torch.Tensor      # [N]
    softness: float = 1000.0
    alpha: float = 100.0
    ray_origin_offset: float = 0.05
    epsilon: float = 1e-12
    max_candidates: int = 64


def pack_distortions(distortions_u: torch.Tensor, distortions_e: torch.Tensor) -> torch.Tensor:
    """Return the interleaved ``[N,R,P,2]`` (u,e) buffer the kernels stream.

    ``Sun.get_distortions`` returns two permuted views of exactly such a buffer
    (``artist/scene/sun.py:227-233``); in that case no copy is made.
    """
    n, r, p = distortions_u.shape
    if (distortions_u.dtype == torch.float32 and distortions_e.dtype == torch.float32
            and distortions_u.stride() == (r * p * 2, p * 2, 2) and distortions_e.stride() == (r * p * 2, p * 2, 2)
            and distortions_e.data_ptr() == distortions_u.data_ptr() + 4
            and distortions_u.storage_offset() % 2 == 0):
        return torch.as_strided(distortions_u, (n, r, p, 2), (r * p * 2, p * 2, 2, 1))
    return torch.stack([distortions_u.float(), distortions_e.float()], dim=-1).contiguous()


@dataclass
class ActivationMap:
    """Index map of ``HeliostatGroup.activate_heliostats`` (``artist/field/heliostat_group.py:256-315``): sample ``k`` is
    heliostat ``rows[k]``; the replicas of heliostat ``s`` are the contiguous samples ``row_start[s] .. row_start[s+1]-1``
    (``repeat_interleave`` order).  Handed to ``trace`` instead of replicated ``[N,P,4]`` surface copies."""

    rows: torch.Tensor        # int32 [N], non-decreasing
    row_start: torch.Tensor   # int32 [Nh+1]

    @staticmethod
    def from_mask(mask: torch.Tensor) -> "ActivationMap":
        counts = mask.long()
        rows = torch.repeat_interleave(torch.arange(counts.numel(), device=mask.device), counts).to(torch.int32)
        row_start = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=mask.device)
        row_start[1:] = torch.cumsum(counts, 0)
        return ActivationMap(rows.contiguous(), row_start)


def replica_sum(per_sample: torch.Tensor, amap: ActivationMap) -> torch.Tensor:
    """``[N,...] -> [Nh,...]``: sum of every heliostat's replica rows (``ab200_replica_sum``, fixed order)."""
    per_sample = _f32(per_sample, "per-sample rows")
    n_src = amap.row_start.numel() - 1
    out = torch.empty(n_src, *per_sample.shape[1:], device=per_sample.device)
    row_elems = per_sample[0].numel() if per_sample.shape[0] else int(torch.tensor(per_sample.shape[1:]).prod())
    _lib.call("ab200_replica_sum", _p(per_sample), _p(amap.row_start), n_src, row_elems, _p(out), _stream())
    return out


_planar_registry: dict = {}   # data_ptr of an interleaved [N,R,P,2] buffer -> (weakref to it, version, planar [2,N,R,P])


def register_planar_distortions(interleaved: torch.Tensor, planar: torch.Tensor) -> None:
    import weakref

    for k in [k for k, v in _planar_registry.items() if v[0]() is None]:
        _planar_registry.pop(k)
    _planar_registry[interleaved.data_ptr()] = (weakref.ref(interleaved), interleaved._version, planar)


def planar_distortions(distortions: torch.Tensor) -> torch.Tensor | None:
    """The de-interleaved ``[2,N,R,P]`` copy of an interleaved ``[N,R,P,2]`` distortion buffer (what the v3 trace kernels
    stream): the one ``Sun.get_distortions`` wrote next to the sample, or - for any other buffer - one made here, once per
    tensor and version (``ab200_deinterleave_distortions``)."""
    import weakref

    if distortions.numel() == 0 or distortions.data_ptr() % 8 != 0:
        return None
    hit = _planar_registry.get(distortions.data_ptr())
    if hit is not None:
        base = hit[0]()
        if base is not None and hit[1] == base._version and hit[2].numel() == distortions.numel() and (
                base is distortions or (base._version == distortions._version and base.shape == distortions.shape)):
            return hit[2]
    n, r, p, _ = distortions.shape
    planar = torch.empty(2, n, r, p, device=distortions.device)
    _lib.call("ab200_deinterleave_distortions", _p(distortions), n * r * p, _p(planar), _stream())
    for k in [k for k, v in _planar_registry.items() if v[0]() is None]:
        _planar_registry.pop(k)
    _planar_registry[distortions.data_ptr()] = (weakref.ref(distortions), distortions._version, planar)
    return planar


def _trace_args(points, normals, incident, distortions, trig, target_idx, targets: TargetTensors, opt: TraceOptions,
                local_rows, flux, intercept, on_target, blocking, dbg=None, blk=None, orientations=None,
                windows=None, src_rows=None) -> _lib.TraceArgs:
    p = points.shape[1]
    n, r = distortions.shape[0], distortions.shape[1]
    a = _lib.TraceArgs()
    a.abi_version = ABI_VERSION
    a.n_samples, a.n_points, a.n_rays = n, p, r
    a.res_e, a.res_u = opt.res_e, opt.res_u
    a.n_local = n if local_rows is None else int(local_rows.numel())
    a.local_rows = _p(local_rows)
    a.points, a.normals, a.incident = _p(points), _p(normals), _p(incident)
    a.distortions, a.trig, a.target_idx = _p(distortions), _p(trig), _p(target_idx)
    a.targets = targets.struct()
    if blk is None:
        a.blockers = _lib.Blockers(0, 0, None, None, None, 1000.0, 100.0, 0.05, 1e-12, 0.0)
    else:
        bi, prims, cand_idx, cand_count = blk
        a.blockers = _lib.Blockers(int(prims.shape[0]), int(bi.max_candidates), _p(prims), _p(cand_idx), _p(cand_count),
                                   float(bi.softness), float(bi.alpha), float(bi.ray_origin_offset), float(bi.epsilon),
                                   6.0 * float(opt.scatter_sigma))
    a.ray_magnitude = float(opt.ray_magnitude)
    a.one_minus_extinction = float(1 - opt.ray_extinction_factor)
    a.reflectivity = float(opt.mirror_reflectivity)
    a.scatter_sigma = float(opt.scatter_sigma)
    a.trig_mode = int(opt.trig_mode)
    a.flags = (_lib.FLAG_FP32_ACCUM if opt.fp32_accumulate else 0) | (_lib.FLAG_ONE_CTA_PER_SAMPLE if opt.one_cta_per_sample else 0)
    a.flux, a.intercept, a.on_target, a.blocking = _p(flux), _p(intercept), _p(on_target), _p(blocking)
    if dbg is not None:
        a.dbg_be, a.dbg_bu, a.dbg_t, a.dbg_lambert = (_p(d) for d in dbg)
    a.stats = _p(trace_stats)
    a.orientations = _p(orientations)
    a.windows = _p(windows)
    a.src_rows = _p(src_rows)
    a.distortions_planar = _p(planar_distortions(distortions)) if (use_planar and dbg is None and trig is None) else None
    return a


# Opt-in (AB200_TRACE_V3=1): hand a de-interleaved distortion copy to the library so that the experimental point-pair
# forward kernel of csrc/trace_v3.cuh runs where it is eligible.  Bit-identical results (tests/test_gpu_trace_parity.py),
# 97 instead of 117 thread-instructions per ray, but at 16 warps per SM it issues less often: 1.09 ms against the general
# kernel's 1.03 ms at the bench size (profiles/r02_trace_v3_experiment.txt) - so it is off by default.
use_planar = os.environ.get("AB200_TRACE_V3", "0") == "1"
trace_stats: torch.Tensor | None = None  # set to a zeroed int64[20] CUDA tensor to collect diagnostics: [0..2] window use
# (threads on the global fallback, window cells, CTAs), [4..10] forward / [12..16] backward phase cycles (tools/phase_stats.py)


last_blocking_overflow: torch.Tensor | None = None  # int32[1]: samples whose candidate list overflowed (must stay 0)
_overflow_pending: list = []   # (pinned host int32[1], event, max_candidates) of earlier calls, checked without a sync
_overflow_pool: list = []      # recycled (pinned flag, event) pairs


def check_blocking_overflow(wait: bool = False) -> None:
    """Raise if a blocking candidate list overflowed ``max_candidates`` (``kMaxBlockCandidates`` = 64 is the hard limit
    of the per-point 64-bit candidate mask): the dropped rectangles would silently be missing from the flux, the
    blocking factors and the gradients, where the reference evaluates every filtered primitive
    (``artist/raytracing/blocking.py:212-354``).  The flag travels to pinned host memory behind the candidate kernel, so
    the check costs no synchronisation: ``trace`` looks at the flags of EARLIER calls that have already arrived
    (``wait=False``); ``HeliostatRayTracer.trace_rays`` callers that need the answer for the current call use ``wait=True``."""
    still = []
    for host, ev, cap in _overflow_pending:
        if wait:
            ev.synchronize()
        if ev.query():
            value = int(host[0])
            _overflow_pool.append((host, ev))
            if value != 0:
                _overflow_pending.clear()
                raise _lib.Ab200Error(
                    f"blocking: {value} heliostat-sample(s) have more than {cap} candidate blockers between them and "
                    f"their target; the candidate list is limited to {cap} (kMaxBlockCandidates = 64) and the flux would "
                    "silently miss the dropped rectangles")
        else:
            still.append((host, ev, cap))
    _overflow_pending[:] = still


def _prepare_blocking(bi: BlockingInputs, opt: TraceOptions, n: int, dev):
    """Pack the primitives and build the per-sample candidate lists (two small kernels)."""
    global last_blocking_overflow
    corners = _f32(bi.corners.detach()[..., :3], "blocking corners")
    spans = _f32(bi.spans.detach()[..., :3], "blocking spans")
    normals = _f32(bi.normals.detach()[..., :3], "blocking normals")
    h = corners.shape[0]
    prims = torch.empty(h, 16, device=dev)
    _lib.call("ab200_blocking_pack", _p(corners), _p(spans), _p(normals), h, float(bi.epsilon), _p(prims), _stream())
    cand_idx = torch.empty(n, bi.max_candidates, dtype=torch.int32, device=dev)
    cand_count = torch.empty(n, dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    owner = _i32(bi.sample_to_blocker, "sample_to_blocker")
    aim = _f32(bi.aim_points.detach(), "aim_points")
    radius = _f32(bi.target_radius.detach(), "target_radius")
    spread = 6.0 * float(opt.scatter_sigma) if opt.scatter_sigma > 0 else 0.02
    _lib.call("ab200_blocking_candidates", _p(prims), h, _p(owner), _p(aim), _p(radius), n, spread, int(bi.max_candidates),
              _p(cand_idx), _p(cand_count), _p(overflow), _stream())
    last_blocking_overflow = overflow
    if torch.cuda.is_current_stream_capturing():
        # CUDA-graph capture: no pinned allocation / event bookkeeping inside the graph; the eager warm-up steps every
        # capture needs have checked this geometry, and `last_blocking_overflow` stays readable after a replay
        return prims, cand_idx, cand_count
    check_blocking_overflow()                      # flags of earlier calls that have arrived by now
    # (pinned flags and events come from a small pool: a pinned allocation per call is a cudaHostAlloc, which serialises
    # with the device - measured as +1.3 ms per step in the end-to-end leg of the motor workload)
    if len(_overflow_pending) >= 64:               # nobody collected for a long time: wait for the oldest
        check_blocking_overflow(wait=True)
    host, ev = _overflow_pool.pop() if _overflow_pool else (torch.empty(1, dtype=torch.int32, pin_memory=True), torch.cuda.Event())
    host.copy_(overflow, non_blocking=True)
    ev.record()
    _overflow_pending.append((host, ev, int(bi.max_candidates)))
    return prims, cand_idx, cand_count


class _TraceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, normals, incident, distortions, trig, target_idx, local_rows, targets, opt, bi, b_corners,
                b_spans, b_normals, orientations, amap):
        points, normals, incident = _f32(points, "points"), _f32(normals, "normals"), _f32(incident, "incident")
        distortions = _f32(distortions, "distortions")
        target_idx_object = target_idx
        target_idx = _i32(target_idx, "target_area_indices")
        if orientations is not None:
            orientations = _f32(orientations, "orientations")
            if orientations.shape != (distortions.shape[0], 4, 4):
                raise _lib.Ab200Error("orientations must be [N,4,4] with one matrix per sample")
        n = distortions.shape[0]
        dev = points.device
        _validate_trace_inputs(points, normals, incident, distortions, trig, target_idx_object, local_rows, targets, amap)
        src_rows = None if amap is None else amap.rows
        flux = torch.empty(n, opt.res_u, opt.res_e, device=dev)
        intercept = torch.empty(n, device=dev)
        on_target = torch.empty(n, device=dev)
        blocking = torch.empty(n, device=dev)
        blk = None
        if bi is not None:
            blk = (bi, *_prepare_blocking(bi, opt, n, dev))
        # the backward re-uses the bitmap windows the forward placed (one per sample)
        windows = torch.empty(n, 4, dtype=torch.int32, device=dev) if any(ctx.needs_input_grad) else None
        args = _trace_args(points, normals, incident, distortions, trig, target_idx, targets, opt, local_rows,
                           flux, intercept, on_target, blocking, blk=blk, orientations=orientations, windows=windows,
                           src_rows=src_rows)
        _lib.call("ab200_trace_fwd", C.byref(args), _stream())
        ctx.amap = amap
        ctx.save_for_backward(points, normals, incident, distortions, trig, target_idx, local_rows, orientations,
                              *(blk[1:] if blk is not None else ()))
        ctx.windows = windows
        ctx.targets, ctx.opt, ctx.bi = targets, opt, bi
        ctx.blocker_shapes = None if bi is None else (b_corners.shape, b_spans.shape, b_normals.shape)
        ctx.mark_non_differentiable(intercept, on_target, blocking)
        return flux, intercept, on_target, blocking

    @staticmethod
    def backward(ctx, g_flux, _gi, _go, _gb):
        saved = ctx.saved_tensors
        points, normals, incident, distortions, trig, target_idx, local_rows, orientations = saved[:8]
        blk = None if ctx.bi is None else (ctx.bi, *saved[8:11])
        if not g_flux.is_cuda:
            raise _lib.Ab200Error("grad_flux must be a CUDA tensor")
        u, e = g_flux.shape[1], g_flux.shape[2]
        if g_flux.dtype == torch.float32 and g_flux.stride(2) == 1 and g_flux.stride(1) == e and g_flux.stride(0) in (0, u * e):
            g_stride = g_flux.stride(0)       # dense, or one [U,E] gradient shared by every sample (expanded view)
        else:
            g_flux = _f32(g_flux, "grad_flux")
            g_stride = u * e
        amap = ctx.amap
        # surface gradients only if somebody wants them (a motor-position optimisation does not: 64 B per point saved)
        need_surface = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        g_points = g_normals = None
        if need_surface:
            shape = (distortions.shape[0], *points.shape[1:])      # per SAMPLE; folded onto the source rows below
            g_points = torch.empty(shape, device=points.device)
            g_normals = torch.empty(shape, device=points.device)
        b = _lib.TraceBwdArgs()
        b.fwd = _trace_args(points, normals, incident, distortions, trig, target_idx, ctx.targets, ctx.opt, local_rows,
                            None, None, None, None, blk=blk, orientations=orientations, windows=ctx.windows,
                            src_rows=None if amap is None else amap.rows)
        b.grad_flux, b.grad_points, b.grad_normals = _p(g_flux), _p(g_points), _p(g_normals)
        b.grad_flux_stride = g_stride
        g_corners = g_spans = g_bnormals = None
        need_blockers = blk is not None and any(ctx.needs_input_grad[10:13]) and os.environ.get("AB200_NO_BLOCKER_GRAD") != "1"
        g_prims = torch.zeros(blk[1].shape[0], 12, device=points.device) if need_blockers else None
        b.grad_prims = _p(g_prims)
        g_scratch = None
        if need_blockers:   # per-CTA rows for the ordered (reproducible) reduction of the blocker gradients
            g_scratch = torch.empty((distortions.shape[0] + 1024) * int(ctx.bi.max_candidates) * 12, device=points.device)
            b.grad_prims_scratch, b.grad_prims_scratch_floats = _p(g_scratch), g_scratch.numel()
        g_ori = None
        if orientations is not None and ctx.needs_input_grad[13]:
            g_ori = torch.zeros_like(orientations)
        b.grad_orientations = _p(g_ori)
        if need_surface or need_blockers or g_ori is not None:
            _lib.call("ab200_trace_bwd", C.byref(b), _stream())
        if need_surface and amap is not None:
            g_points, g_normals = replica_sum(g_points, amap), replica_sum(g_normals, amap)
        if need_blockers:
            cs, ss, ns = ctx.blocker_shapes
            g_corners = torch.zeros(cs, device=points.device)
            g_corners[:, 0, :3] = g_prims[:, 0:3]
            g_spans = torch.zeros(ss, device=points.device)
            g_spans[:, 0, :3] = g_prims[:, 3:6]
            g_spans[:, 1, :3] = g_prims[:, 6:9]
            g_bnormals = torch.zeros(ns, device=points.device)
            g_bnormals[:, :3] = g_prims[:, 9:12]
        return (g_points, g_normals, None, None, None, None, None, None, None, None, g_corners, g_spans, g_bnormals,
                g_ori, None)


def _validate_trace_inputs(points, normals, incident, distortions, trig, target_idx, local_rows, targets, amap=None) -> None:
    """The kernels index raw pointers: every shape they assume is checked here (the reference would raise a shape or
    index error from its eager ops).  Index VALUES are range-checked once per tensor version (cached device read)."""
    if points.dim() != 3 or points.shape[-1] != 4:
        raise _lib.Ab200Error(f"surface points must be [N,P,4], got {tuple(points.shape)}")
    n, p, _ = points.shape
    if normals.shape != points.shape:
        raise _lib.Ab200Error(f"surface normals {tuple(normals.shape)} do not match the points {tuple(points.shape)}")
    if amap is not None:   # points / normals are the group's un-replicated surfaces; N = number of samples of the map
        if amap.rows.dtype != torch.int32 or amap.row_start.dtype != torch.int32 or amap.row_start.numel() != n + 1:
            raise _lib.Ab200Error("activation map: rows / row_start must be int32 with row_start of length Nh+1")
        _check_index_range(amap.rows, n, "activation map rows")
        n = int(amap.rows.numel())
    if distortions.dim() != 4 or distortions.shape[0] != n or distortions.shape[2] != p or distortions.shape[3] != 2:
        raise _lib.Ab200Error(f"distortions must be [N={n},R,P={p},2], got {tuple(distortions.shape)} - was the ray tracer "
                              "built before the heliostats were re-activated?")
    if incident.dim() != 2 or incident.shape[0] != n or incident.shape[1] != 4:
        raise _lib.Ab200Error(f"incident ray directions must be [N={n},4], got {tuple(incident.shape)}")
    if target_idx.numel() != n:
        raise _lib.Ab200Error(f"target_area_indices must have N={n} entries, got {target_idx.numel()}")
    if trig is not None and tuple(trig.shape) != (n, distortions.shape[1], p, 4):
        raise _lib.Ab200Error(f"trig table must be [N,R,P,4], got {tuple(trig.shape)}")
    _check_index_range(target_idx, targets.n_planar + targets.n_cyl, "target_area_indices")
    if local_rows is not None:
        _check_index_range(local_rows, n, "local_rows")


def trace(points, normals, incident, distortions, target_idx, targets: TargetTensors, opt: TraceOptions,
          local_rows: torch.Tensor | None = None, trig: torch.Tensor | None = None,
          blocking: BlockingInputs | None = None, orientations: torch.Tensor | None = None,
          activation: ActivationMap | None = None):
    """Fused forward trace -> ``(flux[N,U,E], intercept[N], on_target[N], blocking[N])``; differentiable
    w.r.t. ``points`` and ``normals`` (and, with ``blocking``, the blockers' corners / spans / normals).

    ``orientations`` ([N,4,4], optional) fuses the alignment ``rows @ O^T`` into the kernels: ``points`` / ``normals``
    are then the UN-aligned rows, the aligned tensors never exist in HBM, and the flux is differentiable w.r.t. the
    orientations too (``ab200_trace_args::orientations``)."""
    if trig is not None:
        trig = _f32(trig, "trig")
    if local_rows is not None:
        local_rows = _i32(local_rows, "local_rows")
    if activation is not None and (trig is not None or opt.trig_mode != _lib.TRIG_POLY or opt.fp32_accumulate):
        # the in-kernel index map exists for the production kernels only (polynomial trig, fixed-point bitmap): the strict /
        # diagnostic modes trace gathered copies, as the reference does (autograd's index_select sums the replicas)
        rows = activation.rows.long()
        points, normals, activation = points.index_select(0, rows), normals.index_select(0, rows), None
    if blocking is None:
        return _TraceFn.apply(points, normals, incident, distortions, trig, target_idx, local_rows, targets, opt, None,
                              None, None, None, orientations, activation)
    return _TraceFn.apply(points, normals, incident, distortions, trig, target_idx, local_rows, targets, opt, blocking,
                          blocking.corners, blocking.spans, blocking.normals, orientations, activation)


def trace_debug(points, normals, incident, distortions, target_idx, targets: TargetTensors, opt: TraceOptions,
                trig: torch.Tensor | None = None):
    """Forward trace that also returns the per-ray ``(be, bu, t, lambert)`` ``[N,R,P]`` parity probes."""
    points, normals, incident = _f32(points, "points"), _f32(normals, "normals"), _f32(incident, "incident")
    distortions = _f32(distortions, "distortions")
    target_idx = _i32(target_idx, "target_area_indices")
    _validate_trace_inputs(points, normals, incident, distortions, trig, target_idx, None, targets)
    n, p, _ = points.shape
    r = distortions.shape[1]
    dev = points.device
    flux = torch.empty(n, opt.res_u, opt.res_e, device=dev)
    fac = [torch.empty(n, device=dev) for _ in range(3)]
    dbg = [torch.zeros(n, r, p, device=dev) for _ in range(4)]
    args = _trace_args(points, normals, incident, distortions, trig, target_idx, targets, opt, None, flux, *fac, dbg=dbg)
    _lib.call("ab200_trace_fwd", C.byref(args), _stream())
    return (flux, *fac), tuple(dbg)


# --------------------------------------------------------------------------------------------
# per-target reduction
# --------------------------------------------------------------------------------------------
class _PerTargetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bitmaps, target_idx, n_targets):
        bitmaps = _f32(bitmaps, "bitmaps_per_heliostat")
        target_idx_object = target_idx
        target_idx = _i32(target_idx, "target_area_indices")
        n, u, e = bitmaps.shape
        if n_targets > 1 and bitmaps.requires_grad:
            prefetch_uniform_target(target_idx_object)
        out = torch.empty(n_targets, u, e, device=bitmaps.device)
        _lib.call("ab200_bitmaps_per_target", _p(bitmaps), _p(target_idx), n, n_targets, u, e, _p(out), _stream())
        ctx.save_for_backward(target_idx)
        ctx.n = n
        ctx.target_idx_object = target_idx_object
        return out

    @staticmethod
    def backward(ctx, g):
        (target_idx,) = ctx.saved_tensors
        # every sample aims at ONE target (the common case): all samples see the same [U,E] gradient - hand an
        # expanded (stride-0) view to the trace backward, which reads it in place instead of an [N,U,E] gather
        t = 0 if g.shape[0] == 1 else _uniform_target(ctx.target_idx_object)
        if t is not None:
            return g[t].expand(ctx.n, -1, -1), None, None
        return g.index_select(0, target_idx.long()), None, None


_uniform_cache: dict = {}


_UNKNOWN = object()


def _index_bounds(idx: torch.Tensor, wait: bool = True):
    """``(min, max)`` of an index tensor, or None if it is empty.  Costs one device->host read per distinct tensor OBJECT
    and version (id + weak reference, so a recycled address or an in-place update is never mistaken for a hit); the
    side-stream check started by ``prefetch_uniform_target`` is collected here if it is on its way.  ``wait=False``
    returns ``_UNKNOWN`` instead of synchronising."""
    import weakref

    hit = _uniform_cache.get(id(idx))
    if hit is not None and hit[0]() is idx and hit[1] == idx._version:
        return hit[2]
    pend = _uniform_pending.get(id(idx))
    if pend is not None and not (pend[0]() is idx and pend[1] == idx._version):
        _uniform_pending.pop(id(idx), None)
        pend = None
    if idx.numel() == 0:
        result = None
    elif pend is not None:
        if not wait and not pend[3].query():
            return _UNKNOWN
        _uniform_pending.pop(id(idx), None)
        pend[3].synchronize()   # long finished by the time anybody asks
        lo_hi = pend[2].tolist()
        result = (int(lo_hi[0]), int(lo_hi[1]))
    else:
        if not wait:
            return _UNKNOWN
        lo, hi = torch.aminmax(idx)
        lo_hi = torch.stack([lo, hi]).tolist()
        result = (int(lo_hi[0]), int(lo_hi[1]))
    if len(_uniform_cache) > 64:
        _uniform_cache.clear()
    _uniform_cache[id(idx)] = (weakref.ref(idx), idx._version, result)
    return result


def _uniform_target(target_idx: torch.Tensor) -> int | None:
    """The single target index all samples share, or None."""
    b = _index_bounds(target_idx)
    return b[0] if b is not None and b[0] == b[1] else None


_seen_index_tensors: dict = {}   # id -> weakref: index tensor OBJECTS that were range-checked synchronously once
_deferred_range_checks: list = []


def _raise_if_out_of_range(b, upper: int, what: str) -> None:
    if b is not None and (b[0] < 0 or b[1] >= upper):
        raise IndexError(f"{what}: values span [{b[0]}, {b[1]}] but must lie in [0, {upper})")


def _check_index_range(idx: torch.Tensor, upper: int, what: str) -> None:
    """Raise ``IndexError`` (as the reference's eager indexing would) if any entry is outside ``[0, upper)``.

    A tensor OBJECT seen for the first time is checked at once (one device->host read, set-up time).  When a known tensor
    was only updated in place (the steady state of an optimisation loop that uploads new indices every step) the check
    runs on a side stream and is collected by a LATER call, so no step ever waits for a device->host read; the kernels
    clamp the indices, so a late error never means an out-of-bounds access happened in between."""
    import weakref

    # collect deferred checks whose result has arrived
    still = []
    for ref, version, up, w in _deferred_range_checks:
        t = ref()
        if t is None or t._version != version:
            continue
        b = _index_bounds(t, wait=False)
        if b is _UNKNOWN:
            still.append((ref, version, up, w))
        else:
            _deferred_range_checks[:] = still
            _raise_if_out_of_range(b, up, w)
    _deferred_range_checks[:] = still
    known = _seen_index_tensors.get(id(idx))
    first_time = known is None or known() is not idx
    b = _index_bounds(idx, wait=first_time)
    if b is _UNKNOWN:
        prefetch_uniform_target(idx)
        _deferred_range_checks.append((weakref.ref(idx), idx._version, upper, what))
    else:
        _raise_if_out_of_range(b, upper, what)
    if first_time:
        if len(_seen_index_tensors) > 256:
            _seen_index_tensors.clear()
        _seen_index_tensors[id(idx)] = weakref.ref(idx)


_uniform_pending: dict = {}
_side_stream: "torch.cuda.Stream | None" = None


def prefetch_uniform_target(target_idx: torch.Tensor) -> None:
    """Start the "do all samples aim at ONE target?" check of ``_uniform_target`` on a side stream, result into pinned host
    memory.  ``trace_rays`` calls this before it launches the forward trace, so by the time the backward pass asks, the
    answer is on the host and no device->host read has to wait in the middle of the step (which would drain the launch
    pipeline).  No-op when the answer for this tensor object and version is already known or on its way."""
    import weakref

    global _side_stream
    if not target_idx.is_cuda or target_idx.numel() == 0:
        return
    k = id(target_idx)
    for table in (_uniform_cache, _uniform_pending):
        hit = table.get(k)
        if hit is not None and hit[0]() is target_idx and hit[1] == target_idx._version:
            return
    dev = target_idx.device
    if _side_stream is None or _side_stream.device != dev:
        _side_stream = torch.cuda.Stream(device=dev)
    _side_stream.wait_stream(torch.cuda.current_stream(dev))
    host = torch.empty(2, dtype=target_idx.dtype, pin_memory=True)
    with torch.cuda.stream(_side_stream):
        lo, hi = torch.aminmax(target_idx)
        host.copy_(torch.stack([lo, hi]), non_blocking=True)
        done = torch.cuda.Event()
        done.record(_side_stream)
    target_idx.record_stream(_side_stream)
    if len(_uniform_pending) > 64:
        _uniform_pending.clear()
    _uniform_pending[k] = (weakref.ref(target_idx), target_idx._version, host, done)


def bitmaps_per_target(bitmaps, target_idx, n_targets: int):
    return _PerTargetFn.apply(bitmaps, target_idx, n_targets)


# --------------------------------------------------------------------------------------------
# flux post-processing: centre of mass, centre-of-mass crop (artist/flux/bitmap.py)
# --------------------------------------------------------------------------------------------
class _FluxCentreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bitmaps, normalised):
        bitmaps = _f32(bitmaps, "bitmaps")
        n, u, e = bitmaps.shape
        moments = torch.empty(n, 3, device=bitmaps.device)
        _lib.call("ab200_flux_moments", _p(bitmaps), n, u, e, 1 if normalised else 0, _p(moments), _stream())
        ctx.save_for_backward(moments)
        ctx.shape, ctx.normalised = (n, u, e), normalised
        return moments[:, 1:].clone()

    @staticmethod
    def backward(ctx, g_centre):
        (moments,) = ctx.saved_tensors
        n, u, e = ctx.shape
        g_centre = _f32(g_centre, "grad_centre")
        g = torch.empty(n, u, e, device=moments.device)
        _lib.call("ab200_flux_moments_bwd", _p(moments), _p(g_centre), n, u, e, 1 if ctx.normalised else 0, _p(g), _stream())
        return g, None


def flux_center_of_mass(bitmaps: torch.Tensor, normalised: bool = False) -> torch.Tensor:
    """``[N,U,E]`` -> ``[N,2]`` (e, u) centre of mass in pixel units, or in the crop's normalised [-1,1] coordinates."""
    return _FluxCentreFn.apply(bitmaps, normalised)


class _FluxCropFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bitmaps, scale):
        bitmaps, scale = _f32(bitmaps, "flux_distributions"), _f32(scale, "crop scale")
        n, u, e = bitmaps.shape
        moments = torch.empty(n, 3, device=bitmaps.device)
        out = torch.empty_like(bitmaps)
        _lib.call("ab200_flux_crop_fwd", _p(bitmaps), _p(scale), n, u, e, _p(moments), _p(out), _stream())
        ctx.save_for_backward(bitmaps, scale, moments)
        return out

    @staticmethod
    def backward(ctx, g_out):
        bitmaps, scale, moments = ctx.saved_tensors
        n, u, e = bitmaps.shape
        g_out = _f32(g_out, "grad_cropped")
        scratch = torch.empty(n, 2, device=bitmaps.device)
        g_in = torch.empty_like(bitmaps)
        _lib.call("ab200_flux_crop_bwd", _p(bitmaps), _p(scale), _p(moments), _p(g_out), n, u, e, _p(scratch), _p(g_in),
                  _stream())
        return g_in, None


def flux_crop_around_center(bitmaps: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """Bilinear crop of every ``[U,E]`` bitmap around its centre of mass; ``scale[n] = (crop_w / target_w, crop_h /
    target_h)``.  Differentiable w.r.t. the bitmaps (through the resampling and through the centre)."""
    return _FluxCropFn.apply(bitmaps, scale)


LOSS_PIXEL, LOSS_KL_DIVERGENCE = 0, 1


class _FluxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, ground_truth, kind):
        prediction = _f32(prediction, "prediction")
        ground_truth = _f32(ground_truth.detach(), "ground_truth")
        if prediction.dim() != 3 or prediction.shape != ground_truth.shape:
            raise _lib.Ab200Error(f"prediction and ground_truth must both be [N,U,E], got {tuple(prediction.shape)} and "
                                  f"{tuple(ground_truth.shape)}")
        n, u, e = prediction.shape
        loss = torch.empty(n, device=prediction.device)
        aux = torch.empty(n, 4, device=prediction.device)
        _lib.call("ab200_flux_loss_fwd", _p(prediction), _p(ground_truth), n, u, e, kind, _p(loss), _p(aux), _stream())
        ctx.save_for_backward(prediction, ground_truth, aux)
        ctx.kind = kind
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        prediction, ground_truth, aux = ctx.saved_tensors
        n, u, e = prediction.shape
        g_loss = _f32(g_loss, "grad_loss")
        g = torch.empty_like(prediction)
        _lib.call("ab200_flux_loss_bwd", _p(prediction), _p(ground_truth), _p(aux), _p(g_loss), n, u, e, ctx.kind, _p(g),
                  _stream())
        return g, None, None


def flux_loss(prediction: torch.Tensor, ground_truth: torch.Tensor, kind: int) -> torch.Tensor:
    """Per-sample loss ``[N]`` of ``[N,U,E]`` flux bitmaps against a (constant) ground truth, reduced over the whole
    bitmap: ``LOSS_PIXEL`` (``artist/optim/loss.py:251-319``) or ``LOSS_KL_DIVERGENCE`` (``:322-410``).  One fused
    kernel forward, one backward (gradient w.r.t. the prediction)."""
    return _FluxLossFn.apply(prediction, ground_truth, kind)


# --------------------------------------------------------------------------------------------
# NURBS
# --------------------------------------------------------------------------------------------
_grid_cache: dict = {}


def detect_evaluation_grid(eval_points: torch.Tensor) -> tuple[int, int]:
    """``(pu, pv)`` if the ``[N,F,K,2]`` evaluation points are ONE sorted cartesian grid ``u_i x v_j`` (v fastest)
    shared by all surfaces and facets - what ``create_nurbs_evaluation_grid(...).expand(N, F, -1, -1)`` gives -
    else ``(0, 0)``.  One small device->host copy per distinct tensor; the answer is cached."""
    key = (eval_points.data_ptr(), tuple(eval_points.shape), tuple(eval_points.stride()), eval_points._version)
    hit = _grid_cache.get(key)
    if hit is not None:
        return hit
    result = (0, 0)
    n, f, k, _ = eval_points.shape
    shared = (n == 1 or eval_points.stride(0) == 0) and (f == 1 or eval_points.stride(1) == 0)
    if not shared:  # materialised copies of one grid are fine too
        shared = bool((eval_points == eval_points[:1, :1]).all().item())
    if shared and k >= 4:
        base = eval_points[0, 0].detach().cpu()
        u, v = base[:, 0], base[:, 1]
        pv = int((u == u[0]).sum())
        if 1 < pv < k and k % pv == 0:
            pu = k // pv
            ug, vg = u.reshape(pu, pv), v.reshape(pu, pv)
            ok = bool((ug == ug[:, :1]).all() and (vg == vg[:1]).all() and (ug[1:, 0] > ug[:-1, 0]).all()
                      and (vg[0, 1:] > vg[0, :-1]).all())
            if ok and pu <= 128 and pv <= 128:
                result = (pu, pv)
    if len(_grid_cache) > 64:
        _grid_cache.clear()
    _grid_cache[key] = result
    return result


def _nurbs_args(cp, eval_points, knots_u, knots_v, degree_u, degree_v, canting, translations, points, normals):
    n, f, cu, cv, _ = cp.shape
    a = _lib.NurbsArgs()
    a.abi_version = ABI_VERSION
    a.n_surfaces, a.n_facets = n, f
    a.n_ctrl_u, a.n_ctrl_v, a.degree_u, a.degree_v = cu, cv, degree_u, degree_v
    a.control_points = _p(cp)
    # evaluation points [N,F,K,2] possibly an expanded view of a shared [K,2] grid
    a.n_eval = eval_points.shape[2]
    st = eval_points.stride()
    if st[2] != 2 or st[3] != 1:
        raise _lib.Ab200Error("evaluation_points must be (u,v)-interleaved along the last two dims")
    a.eval_points = eval_points.data_ptr()
    a.eval_stride_n = st[0] if eval_points.shape[0] > 1 else 0
    a.eval_stride_f = st[1] if eval_points.shape[1] > 1 else 0
    a.knots_u, a.knots_v = _p(knots_u), _p(knots_v)
    a.canting, a.facet_translations = _p(canting), _p(translations)
    a.grid_u, a.grid_v = detect_evaluation_grid(eval_points)
    a.points, a.normals = _p(points), _p(normals)
    return a


class _NurbsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cp, eval_points, knots_u, knots_v, degree_u, degree_v, canting, translations):
        cp = _f32(cp, "control_points")
        if eval_points.dtype != torch.float32 or not eval_points.is_cuda:
            raise _lib.Ab200Error("evaluation_points must be a CUDA float32 tensor")
        if eval_points.stride()[-1] != 1 or eval_points.stride()[-2] != 2:
            eval_points = eval_points.contiguous()
        n, f = cp.shape[:2]
        k = eval_points.shape[2]
        points = torch.empty(n, f, k, 4, device=cp.device)
        normals = torch.empty(n, f, k, 4, device=cp.device)
        canting = None if canting is None else _f32(canting, "canting")
        translations = None if translations is None else _f32(translations, "facet_translations")
        args = _nurbs_args(cp, eval_points, knots_u, knots_v, degree_u, degree_v, canting, translations, points, normals)
        _lib.call("ab200_nurbs_fwd", C.byref(args), _stream())
        ctx.save_for_backward(cp, eval_points, knots_u, knots_v, canting, translations)
        ctx.deg = (degree_u, degree_v)
        return points, normals

    @staticmethod
    def backward(ctx, g_points, g_normals):
        cp, eval_points, knots_u, knots_v, canting, translations = ctx.saved_tensors
        g_points = _f32(g_points, "grad_points")
        g_normals = _f32(g_normals, "grad_normals")
        g_cp = torch.empty_like(cp)
        b = _lib.NurbsBwdArgs()
        b.fwd = _nurbs_args(cp, eval_points, knots_u, knots_v, ctx.deg[0], ctx.deg[1], canting, translations, None, None)
        b.grad_points, b.grad_normals, b.grad_control_points = _p(g_points), _p(g_normals), _p(g_cp)
        _lib.call("ab200_nurbs_bwd", C.byref(b), _stream())
        return g_cp, None, None, None, None, None, None, None


def nurbs_points_and_normals(control_points, evaluation_points, knots_u, knots_v, degree_u: int, degree_v: int,
                             canting=None, facet_translations=None):
    return _NurbsFn.apply(control_points, evaluation_points, knots_u, knots_v, degree_u, degree_v, canting,
                          facet_translations)


# --------------------------------------------------------------------------------------------
# kinematics
# --------------------------------------------------------------------------------------------
def _actuator_rows(act_non_opt: torch.Tensor) -> torch.Tensor:
    """The kernels index the non-optimizable actuator parameters as ``[N,7,2]``; ideal actuators loaded from a
    scenario file carry only the first four rows (``h5_scenario_parser.py:574-640``) - pad them with zeros."""
    if act_non_opt.dim() == 3 and act_non_opt.shape[1] < 7:
        act_non_opt = torch.nn.functional.pad(act_non_opt, (0, 0, 0, 7 - act_non_opt.shape[1]))
    return act_non_opt


def _kin_args(positions, trans_dev, rot_dev, act_non_opt, act_opt, offset, linear: bool):
    k = _lib.KinematicsArgs()
    k.abi_version = ABI_VERSION
    k.n = positions.shape[0]
    k.linear_actuators = 1 if linear else 0
    k.positions, k.translation_dev, k.rotation_dev = _p(positions), _p(trans_dev), _p(rot_dev)
    k.actuator_non_opt, k.actuator_opt, k.orientation_offset = _p(act_non_opt), _p(act_opt), _p(offset)
    return k


class _KinematicsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, motor, rot_dev, trans_dev, act_opt, positions, act_non_opt, offset, linear):
        motor, rot_dev, trans_dev = _f32(motor, "motor_positions"), _f32(rot_dev, "rotation_deviation"), _f32(trans_dev, "translation_deviation")
        positions, act_non_opt, offset = _f32(positions, "positions"), _f32(_actuator_rows(act_non_opt), "actuator params"), _f32(offset, "offset")
        has_opt = act_opt is not None and act_opt.numel() > 0
        act_opt_c = _f32(act_opt, "actuator_opt") if has_opt else None
        if linear and not has_opt:
            raise _lib.Ab200Error("linear actuators need optimizable parameters [N,2,2]")
        out = torch.empty(motor.shape[0], 4, 4, device=motor.device)
        k = _kin_args(positions, trans_dev, rot_dev, act_non_opt, act_opt_c, offset, linear)
        _lib.call("ab200_kinematics_fwd", C.byref(k), _p(motor), _p(out), _stream())
        ctx.save_for_backward(motor, rot_dev, trans_dev, act_opt_c, positions, act_non_opt, offset)
        ctx.linear = linear
        return out

    @staticmethod
    def backward(ctx, g_out):
        motor, rot_dev, trans_dev, act_opt, positions, act_non_opt, offset = ctx.saved_tensors
        g_out = _f32(g_out, "grad_orientations")
        need = ctx.needs_input_grad
        g_motor = torch.empty_like(motor) if need[0] else None
        g_rot = torch.empty_like(rot_dev) if need[1] else None
        g_trans = torch.empty_like(trans_dev) if need[2] else None
        g_act = torch.empty_like(act_opt) if (need[3] and act_opt is not None) else None
        g_pos = torch.empty_like(positions) if need[4] else None
        k = _kin_args(positions, trans_dev, rot_dev, act_non_opt, act_opt, offset, ctx.linear)
        _lib.call("ab200_kinematics_bwd", C.byref(k), _p(motor), _p(g_out), _p(g_motor), _p(g_rot), _p(g_trans),
                  _p(g_act), _p(g_pos), _stream())
        return g_motor, g_rot, g_trans, g_act, g_pos, None, None, None


def kinematics_orientations(motor, rot_dev, trans_dev, act_opt, positions, act_non_opt, offset, linear: bool):
    return _KinematicsFn.apply(motor, rot_dev, trans_dev, act_opt, positions, act_non_opt, offset, linear)


def kinematics_align_incident(incident, aim_points, rot_dev, trans_dev, act_opt, positions, act_non_opt, offset,
                              linear: bool, max_iterations: int = 4, min_eps: float = 1e-4):
    """Forward-only: the orientations carry no autograd graph.  The reference's loop IS differentiable w.r.t. the kinematic
    deviations and actuator parameters (``kinematics_rigid_body.py:540-634``) although none of its callers uses that; a
    caller who does gets a warning here instead of silently missing gradients - ``kinematics_orientations`` on the
    converged motor positions (``RigidBody.motor_positions_to_orientations``) is the differentiable route."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (rot_dev, trans_dev, act_opt)):
        warn_no_graph("incident_ray_directions_to_orientations",
                      "align with motor_positions_to_orientations(kinematics.active_motor_positions) for gradients")
    with torch.no_grad():
        incident, aim_points = _f32(incident, "incident"), _f32(aim_points, "aim_points")
        rot_dev, trans_dev = _f32(rot_dev.detach(), "rotation_deviation"), _f32(trans_dev.detach(), "translation_deviation")
        positions, act_non_opt, offset = _f32(positions, "positions"), _f32(_actuator_rows(act_non_opt), "actuator params"), _f32(offset, "offset")
        has_opt = act_opt is not None and act_opt.numel() > 0
        act_opt_c = _f32(act_opt.detach(), "actuator_opt") if has_opt else None
        n = incident.shape[0]
        out = torch.empty(n, 4, 4, device=incident.device)
        motor = torch.empty(n, 2, device=incident.device)
        scratch = torch.empty(4 * n + 8, device=incident.device)
        k = _kin_args(positions, trans_dev, rot_dev, act_non_opt, act_opt_c, offset, linear)
        _lib.call("ab200_kinematics_align_incident", C.byref(k), _p(incident), _p(aim_points), int(max_iterations),
                  float(min_eps), _p(out), _p(motor), _p(scratch), _stream())
    return out, motor


# --------------------------------------------------------------------------------------------
# apply orientation
# --------------------------------------------------------------------------------------------
class _AlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, normals, orientations, src_row):
        points, normals, orientations = _f32(points, "surface_points"), _f32(normals, "surface_normals"), _f32(orientations, "orientations")
        n = orientations.shape[0]
        p = points.shape[1]
        out_p = torch.empty(n, p, 4, device=points.device)
        out_n = torch.empty(n, p, 4, device=points.device)
        _lib.call("ab200_align_fwd", _p(points), _p(normals), _p(orientations), _p(src_row), n, p, _p(out_p), _p(out_n),
                  _stream())
        ctx.save_for_backward(points, normals, orientations, src_row)
        return out_p, out_n

    @staticmethod
    def backward(ctx, g_op, g_on):
        points, normals, orientations, src_row = ctx.saved_tensors
        n, p = orientations.shape[0], points.shape[1]
        need_data = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_ori = ctx.needs_input_grad[2]
        g_op = _f32(g_op, "grad_points")
        g_on = _f32(g_on, "grad_normals")
        g_p = torch.empty(n, p, 4, device=points.device) if need_data else None
        g_n = torch.empty(n, p, 4, device=points.device) if need_data else None
        g_o = torch.empty(n, 4, 4, device=points.device) if need_ori else None
        _lib.call("ab200_align_bwd", _p(points), _p(normals), _p(orientations), _p(src_row), n, p, _p(g_op), _p(g_on),
                  _p(g_p), _p(g_n), _p(g_o), _stream())
        if need_data and src_row is not None:
            # replicated samples share one source row: sum their gradients (repeat_interleave backward)
            gp_src = torch.zeros_like(points).index_add_(0, src_row.long(), g_p)
            gn_src = torch.zeros_like(normals).index_add_(0, src_row.long(), g_n)
            g_p, g_n = gp_src, gn_src
        return g_p, g_n, g_o, None


def align_surfaces(points, normals, orientations, src_row: torch.Tensor | None = None):
    """``points @ O^T, normals @ O^T``; ``src_row`` folds ``repeat_interleave`` activation into the pass."""
    if src_row is not None:
        src_row = _i32(src_row, "src_row")
    return _AlignFn.apply(points, normals, orientations, src_row)


def debug_trig(angles: torch.Tensor, mode: int):
    angles = _f32(angles, "angles").reshape(-1)
    s = torch.empty_like(angles)
    c = torch.empty_like(angles)
    _lib.call("ab200_debug_trig", _p(angles), angles.numel(), int(mode), _p(s), _p(c), _stream())
    return s, c


def debug_const_div(a: torch.Tensor, b: float):
    a = _f32(a, "a").reshape(-1)
    qf, qi = torch.empty_like(a), torch.empty_like(a)
    _lib.call("ab200_debug_const_div", _p(a), a.numel(), float(b), _p(qf), _p(qi), _stream())
    return qf, qi


def debug_div_regular(a: torch.Tensor, b: torch.Tensor):
    a, b = _f32(a, "a").reshape(-1), _f32(b, "b").reshape(-1)
    qf, qi = torch.empty_like(a), torch.empty_like(a)
    _lib.call("ab200_debug_div_regular", _p(a), _p(b), a.numel(), _p(qf), _p(qi), _stream())
    return qf, qi
