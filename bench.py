#!/usr/bin/env python
"""Benchmark of the heliostat ray-tracing hot path (BASELINE.json metric: rays/s forward+backward).

One "step" = one full forward+backward flux prediction of a synthetic Juelich-scale field, through the
reference-shaped class API: NURBS control points -> surface points/normals -> alignment to the sun ->
fused trace (reflect, scatter, intersect, splat) -> per-target flux [-> NCCL all-reduce over ranks] ->
loss -> backward to the control points -> Adam update.  rays = N * P * R per step and rank.

  python bench.py --gpus N --steps K --warmup W          own arm (this repo's CUDA path), one rank per GPU
  python bench.py --impl reference ...                   reference arm: the CPU oracle port of the reference's
                                                         eager-PyTorch path on the box's host cores (bounded sample)

Prints ONE JSON line (rank 0).  Contract details: see the task description / DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# ---- workload (SURVEY.md 8d, config 3/4 of BASELINE.json) ------------------------------------------
N_HELIOSTATS = 2048          # heliostat-samples per GPU ("~2000-heliostat field"), weak scaling
POINTS_PER_FACET = (50, 50)  # 4 facets -> P = 10 000 surface points
CONTROL_POINTS = (10, 10)
RAYS = 10                    # rays per surface point
RES = (256, 256)             # bitmap E x U
CPU_SAMPLE_HELIOSTATS = 64   # bounded CPU sample of the same per-heliostat workload (64-256)
SURFACE_BUMP = float(os.environ.get("AB200_BENCH_BUMP", "1e-4"))  # control-point height noise (m): 0.1 mm ~ 0.6 mrad slope error


def bytes_per_ray(r: int, p: int, ue: int) -> tuple[float, float]:
    """Algorithmic HBM bytes per ray of the trace kernels: forward, backward."""
    return 8 + 32 / r + 4 * ue / (r * p), 8 + 64 / r + 4 * ue / (r * p)


def measured_traffic(kernel: str, rays_per_launch: int) -> float | None:
    """DRAM bytes per launch of `kernel` from the committed ``ncu --set full`` capture (profiles/trace_traffic.json,
    written by tools/ncu_traffic.py) - only if it was taken on the same workload size; else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "trace_traffic.json")) as fh:
            t = json.load(fh)
        if int(t["rays_per_launch"]) != int(rays_per_launch):
            return None
        return float(t["kernels"][kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def issue_roof(kernel: str, rays_per_launch: int, kernel_ms: float, sm_mhz: float | None) -> dict | None:
    """The roof that actually binds the trace kernels: warp-instruction issue (4 per clock and SM).  Instruction count per
    launch from the committed ncu capture (profiles/trace_traffic.json), duration and SM clock measured live."""
    try:
        with open(os.path.join(ROOT, "profiles", "trace_traffic.json")) as fh:
            t = json.load(fh)
        if int(t["rays_per_launch"]) != int(rays_per_launch):
            return None
        inst = float(t["kernels"][kernel]["warp_instructions_per_launch"])
    except Exception:
        return None
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clock_hz = (sm_mhz or 1965.0) * 1e6
    peak = sms * 4 * clock_hz
    achieved = inst / (kernel_ms * 1e-3)
    return {"bound": "issue", "kernel": kernel, "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G warp-instructions/s",
            "frac": achieved / peak, "warp_instructions_per_launch": inst,
            "note": "smsp__inst_executed.sum of the committed ncu capture / live kernel time; peak = SMs x 4 issue slots x live SM clock"}


def measured_peak_hbm() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed region (NVML, 20 ms period, own thread)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int) -> None:
        self.index, self.sm, self.bits, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self._err = None

    def _visible_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self) -> None:
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def run():
                while not self._stop.is_set():
                    try:
                        self.sm.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.bits |= int(get_reasons(h))
                    except Exception as exc:  # pragma: no cover
                        self._err = str(exc)
                        return
                    time.sleep(0.02)

            self._thread = threading.Thread(target=run, daemon=True)
            self._thread.start()
        except Exception as exc:
            self._err = str(exc)

    def stop(self) -> dict:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        sm = sorted(self.sm)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": len(sm)}
        if self._err:
            out["error"] = self._err
        return out


# ---- own arm ---------------------------------------------------------------------------------------
class Workload:
    """Device-resident synthetic field + the fwd/bwd step through the public class API.

    kind "surface" (headline; BASELINE.json configs 3/4): parameters = NURBS control points, planar target, blocking off.
    kind "motor" (configs[4], tutorial 05 / AimPointOptimizer shape): parameters = motor positions, the tilted cylindrical
    receiver, blocking ON, gradients through the trace's orientation gradient and the kinematics backward.
    strong=True: ONE field shared by all ranks, split by the reference's sampler contract (HeliostatRayTracer(world_size,
    rank), artist/raytracing/sampling.py:129-146); everything before the trace is replicated per rank as in the reference."""

    def __init__(self, dev: torch.device, n: int, world: int, rank: int, kind: str = "surface", strong: bool = False) -> None:
        from artist_b200 import HeliostatRayTracer, NURBSSurfaces, build_synthetic_scenario
        from artist_b200.nurbs import create_nurbs_evaluation_grid

        self.dev, self.n, self.world, self.kind, self.strong = dev, n, world, kind, strong
        field_seed = 0 if strong else rank
        self.scenario, self.group = build_synthetic_scenario(
            n, number_of_rays=RAYS, points_per_facet=POINTS_PER_FACET, control_points=CONTROL_POINTS,
            surface_bump=SURFACE_BUMP, seed=field_seed, device=dev)
        g = self.group
        # index 1 = the cylindrical receiver of the synthetic tower (AB200_BENCH_MOTOR_TARGET=0: planar, diagnosis only)
        target = int(os.environ.get("AB200_BENCH_MOTOR_TARGET", "1")) if kind == "motor" else None
        self.mask, self.tidx, self.inc = self.scenario.index_mapping(g, single_target_area_index=target or 0)
        self.inc = self.inc.contiguous()
        self.aim = self.scenario.solar_tower.get_centers_of_target_areas(self.tidx)
        g.activate_heliostats(self.mask)
        if kind == "surface":
            self.param = g.nurbs_control_points.detach().clone().requires_grad_(True)
            g.nurbs_control_points = self.param
            g.activate_heliostats(self.mask)
            self.surf = NURBSSurfaces(g.nurbs_degrees, self.param, device=dev)
            grid = create_nurbs_evaluation_grid(torch.tensor(POINTS_PER_FACET), device=dev)
            self.ev = grid[None, None].expand(n, g.number_of_facets_per_heliostat, -1, -1)
        g.align_surfaces_with_incident_ray_directions(self.aim, self.inc, self.mask)
        if kind == "motor":
            self.param = g.kinematics.active_motor_positions.detach().clone().requires_grad_(True)
        tracer_seed = 7 if strong else 7 + rank
        self.tracer = HeliostatRayTracer(self.scenario, g, blocking_active=(kind == "motor"), random_seed=tracer_seed,
                                         world_size=world if strong else 1, rank=rank if strong else 0,
                                         bitmap_resolution=torch.tensor(RES))
        if os.environ.get("AB200_FORCE_ONE_CTA") == "1":     # tuning: the one-CTA-per-sample kernels whatever the sample count
            self.tracer._force_one_cta_per_sample = True
        self.opt = torch.optim.Adam([self.param], lr=1e-6 if kind == "surface" else 1e-3, fused=True)
        self.p = g.surface_points.shape[1]
        self.n_local = len(self.tracer.distortions_sampler.rank_indices)
        self.rays_per_step = self.n_local * self.p * RAYS
        # pinned host mirrors + device staging of the end-to-end leg (double-buffered, see step_e2e)
        self.h_param = self.param.detach().cpu().pin_memory()
        self.h_inc = self.inc.cpu().pin_memory()
        self.h_tidx = self.tidx.cpu().pin_memory()
        n_t = int(self.scenario.solar_tower.number_of_target_areas_per_type.sum())
        self.h_flux = [torch.empty(n_t, RES[1], RES[0]).pin_memory() for _ in range(2)]
        self.h_grad = [torch.empty_like(self.h_param).pin_memory() for _ in range(2)]
        self.h_loss = [torch.empty(1).pin_memory() for _ in range(2)]
        self.d_stage = [(torch.empty_like(self.param.detach()), torch.empty_like(self.inc), torch.empty_like(self.tidx))
                        for _ in range(2)]
        self.s_h2d, self.s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.ev_up = [torch.cuda.Event() for _ in range(2)]
        self.ev_used = [torch.cuda.Event() for _ in range(2)]
        self.ev_down = [torch.cuda.Event() for _ in range(2)]
        self.k = 0

    def forward_local(self) -> torch.Tensor:
        """This rank's per-target flux [T,U,E] (before the sum over ranks), with the autograd graph."""
        g = self.group
        g.activate_heliostats(self.mask)
        if self.kind == "surface":
            pts, nrm = self.surf.calculate_surface_points_and_normals(self.ev, g.active_canting, g.active_facet_translations)
            g.active_surface_points = pts.reshape(self.n, -1, 4)
            g.active_surface_normals = nrm.reshape(self.n, -1, 4)
            g.align_surfaces_with_incident_ray_directions(self.aim, self.inc, self.mask)
        else:
            g.align_surfaces_with_motor_positions(self.param, self.mask)
        flux, _, _, _ = self.tracer.trace_rays(self.inc, self.mask, self.tidx)
        return self.tracer.get_bitmaps_per_target(flux, self.tidx)

    def step(self) -> torch.Tensor:
        self.opt.zero_grad(set_to_none=True)
        total = self.forward_local()
        if self.world > 1:
            import torch.distributed.nn.functional as dist_fn

            total = dist_fn.all_reduce(total)  # autograd-aware SUM over ranks (NCCL, NVLink)
        loss = (total * total).mean()
        loss.backward()
        if self.strong and self.world > 1:
            # replicated parameters, each rank holds the gradient of its own samples (SurfaceReconstructor reduces the
            # same way, artist/optim/surface_reconstructor.py:766-777)
            torch.distributed.all_reduce(self.param.grad)
            self.param.grad /= self.world
        self.opt.step()
        self.last_total = total
        return loss

    def step_e2e(self) -> None:
        """Same step with HOST buffers.  Every step uploads its parameters / incident directions / target indices from
        pinned memory and downloads its per-target flux, loss and parameter gradient - on two copy streams, so that the
        upload of step i and the download of step i-1 overlap the kernels of the step in flight (device staging buffers
        and host result buffers are double-buffered; events order every re-use).  The host holds the results of step i-1
        when step_e2e(i) returns - the one-step lag of a pipelined optimisation loop; flush_e2e() collects the last."""
        b = self.k & 1
        cs = torch.cuda.current_stream()
        with torch.cuda.stream(self.s_h2d):
            self.s_h2d.wait_event(self.ev_used[b])          # staging buffer b was consumed two steps ago
            sp, si, st = self.d_stage[b]
            sp.copy_(self.h_param, non_blocking=True)
            si.copy_(self.h_inc, non_blocking=True)
            st.copy_(self.h_tidx, non_blocking=True)
            self.ev_up[b].record(self.s_h2d)
        cs.wait_event(self.ev_up[b])
        with torch.no_grad():
            self.param.copy_(sp)
            self.inc.copy_(si)
            self.tidx.copy_(st)
        self.ev_used[b].record(cs)
        loss = self.step()
        done = torch.cuda.Event()
        done.record(cs)
        total, grad, loss = self.last_total.detach(), self.param.grad, loss.detach().reshape(1)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(done)
            self.h_flux[b].copy_(total, non_blocking=True)
            self.h_grad[b].copy_(grad, non_blocking=True)
            self.h_loss[b].copy_(loss, non_blocking=True)
            for t in (total, grad, loss):
                t.record_stream(self.s_d2h)
            self.ev_down[b].record(self.s_d2h)
        if self.k > 0:
            self.ev_down[b ^ 1].synchronize()               # results of the previous step are on the host now
        self.k += 1

    def flush_e2e(self) -> None:
        self.ev_down[(self.k - 1) & 1].synchronize()
        self.s_h2d.synchronize()

    @property
    def e2e_bytes(self) -> tuple[int, int]:
        h2d = self.h_param.numel() * 4 + self.h_inc.numel() * 4 + self.h_tidx.numel() * 4
        d2h = self.h_flux[0].numel() * 4 + self.h_grad[0].numel() * 4 + 4
        return h2d, d2h


def timed_loop(fn, steps: int, dev: torch.device, distributed: bool, finish=None) -> float:
    """ms per step over exactly `steps` steps: barrier + synchronize on both sides, CUDA events, max over ranks."""
    if distributed:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if finish is not None:
        finish()     # (end-to-end leg: the last step's download on the copy stream has reached the host)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if distributed:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.barrier()
        ms = float(t.item())
    return ms / steps


def cpu_oracle_step_factory(n: int, threads: int, kind: str = "surface"):
    """The oracle port of the reference's eager path, same per-heliostat workload, on the host cores."""
    from artist_b200.scenario.synthetic import synthetic_field_tensors
    from oracle import artist_oracle as O

    torch.set_num_threads(threads)
    ft = synthetic_field_tensors(n, control_points=CONTROL_POINTS, surface_bump=SURFACE_BUMP, seed=0)
    tg = O.targets_from_field_tensors(ft)
    ev = O.nurbs_evaluation_grid(*POINTS_PER_FACET)[None, None].expand(n, 4, -1, -1)
    tidx = torch.full((n,), 1 if kind == "motor" else 0, dtype=torch.int32)
    inc = torch.tensor([0.0, 1.0, 0.0, 0.0]).expand(n, -1).contiguous()
    aim = O.aim_points(tg, tidx)
    kin = O.Kin(ft["positions"], ft["translation_deviations"], ft["rotation_deviations"],
                ft["actuator_non_optimizable"], ft["actuator_optimizable"], True)
    p = 4 * POINTS_PER_FACET[0] * POINTS_PER_FACET[1]
    du, de = O.sun_distortions(RAYS, p, n, 7)
    first = {}
    if kind == "surface":
        param = ft["nurbs_control_points"].clone().requires_grad_(True)
        opt = torch.optim.Adam([param], lr=1e-6)
    else:
        with torch.no_grad():
            pts0, nrm0 = O.nurbs_points_and_normals(ft["nurbs_control_points"], 3, 3, ev, ft["canting"], ft["facet_translations"])
            pts0, nrm0 = pts0.reshape(n, -1, 4), nrm0.reshape(n, -1, 4)
            _, motor0 = O.incident_ray_directions_to_orientations(kin, inc, aim)
        param = motor0.clone().requires_grad_(True)
        opt = torch.optim.Adam([param], lr=1e-3)
        owner = torch.arange(n)

    def step():
        opt.zero_grad(set_to_none=True)
        if kind == "surface":
            pts, nrm = O.nurbs_points_and_normals(param, 3, 3, ev, ft["canting"], ft["facet_translations"])
            with torch.no_grad():
                ori, _ = O.incident_ray_directions_to_orientations(kin, inc, aim)
            ap, an = O.align_surfaces(pts.reshape(n, -1, 4), nrm.reshape(n, -1, 4), ori)
            blocking = None
        else:
            ori = O.motor_positions_to_orientations(kin, param)
            ap, an = O.align_surfaces(pts0, nrm0, ori)
            corners, spans, bn = O.blocking_primitives(ap)
            blocking = dict(corners=corners, spans=spans, normals=bn, sample_to_blocker=owner)
        flux, *_ = O.trace_rays(ap, an, inc, du, de, tidx, tg, RES, batch_size=min(64, n), blocking=blocking)
        total = O.bitmaps_per_target(flux, tidx, tg.n_total)
        loss = (total * total).mean()
        loss.backward()
        if not first:   # outputs of the very first step (initial parameters): the parity reference of the own arm
            first.update(total=total.detach().clone(), grad=param.grad.detach().clone(), du=du, de=de)
        opt.step()
        return loss

    step.first = first
    return step, n * p * RAYS


def parity_against_cpu_step(dev: torch.device, first: dict, kind: str, first_crt: dict | None) -> dict:
    """The own arm on the CPU baseline's sample (same field, same distortion samples, first step): per-target flux and
    parameter gradient against the oracle's - the 'flux max rel err' half of BASELINE.json's metric.  The GPU side is forced
    onto the one-CTA-per-sample kernels the full field runs (AB200_FLAG_ONE_CTA_PER_SAMPLE), not the split-mode kernels a
    48-heliostat slice would otherwise pick.  `first_crt` = the same oracle step with correctly rounded cos / sin
    (oracle.correctly_rounded_trig): the reference's own sensitivity to the last bit of torch-CPU trig, reported beside."""
    from artist_b200 import ops

    n = first["du"].shape[0]
    wl = Workload(dev, n, 1, 0, kind=kind)
    wl.tracer._force_one_cta_per_sample = True
    ds = wl.tracer.distortions_dataset
    ds.distortions_u, ds.distortions_e = first["du"].to(dev), first["de"].to(dev)
    wl.tracer._packed = ops.pack_distortions(ds.distortions_u, ds.distortions_e)
    ops.trace_stats = torch.zeros(20, dtype=torch.int64, device=dev)
    wl.step()
    one_cta = int(ops.trace_stats[2]) == n
    ops.trace_stats = None
    total, grad = wl.last_total.detach().cpu(), wl.param.grad.detach().cpu()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    out = {"flux_max_rel_err": rel(total, first["total"]), "grad_max_rel_err": rel(grad, first["grad"]),
           "kernels": "one-CTA-per-sample (as the full field)" if one_cta else "split mode",
           "sample": f"{n} heliostats of the bench field, first step, distortion samples of the CPU baseline step; "
                     "per-target flux relative to its peak, parameter gradient relative to its largest entry; "
                     "CPU side = oracle port (pinned bit-exact to the reference), GPU side = device trig + fixed-point bitmap"}
    if first_crt:
        out["vs_oracle_with_correctly_rounded_trig"] = {
            "flux_max_rel_err": rel(total, first_crt["total"]), "grad_max_rel_err": rel(grad, first_crt["grad"]),
            "oracle_own_shift": {"flux": rel(first_crt["total"], first["total"]), "grad": rel(first_crt["grad"], first["grad"])},
            "note": "torch-CPU cos (SLEEF) is not correctly rounded for ~8.6 % of sun-shape angles; the kernels' polynomial "
                    "reproduces torch's value for 99.94 % of them (so the plain comparison above is the one that counts); "
                    "oracle_own_shift = how far that last bit alone moves the ORACLE"}
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step, rays = cpu_oracle_step_factory(CPU_SAMPLE_HELIOSTATS, threads, args.workload)
    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = rays / dt
    sample = (f"{CPU_SAMPLE_HELIOSTATS} heliostats x {rays // CPU_SAMPLE_HELIOSTATS} rays of the same synthetic field per step "
              f"(the full {N_HELIOSTATS}-heliostat step needs ~0.6 KB/ray of autograd state on the CPU); throughput scales "
              "linearly in heliostats")
    out = {
        "impl": "reference", "metric": "rays/s forward+backward", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.workload, args.scaling),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def workload_config(gpus: int, kind: str = "surface", scaling: str = "weak") -> dict:
    per = "per GPU" if scaling == "weak" else f"in ONE field split over {gpus} GPU(s) by the sampler contract"
    if kind == "surface":
        what = ("planar 8x8 m target, blocking off; step = NURBS eval + alignment + fused trace + per-target flux + loss + "
                "backward to control points + Adam")
    else:
        what = ("tilted cylindrical receiver (r 4.14 m), blocking ON; step = motor positions -> orientations + fused trace + "
                "per-target flux + loss + backward to the motor positions + Adam (BASELINE.json configs[4])")
    cfg = {
        "workload": (f"synthetic Juelich-scale field, {N_HELIOSTATS} heliostats {per} x 4 facets x "
                     f"{POINTS_PER_FACET[0]}x{POINTS_PER_FACET[1]} surface points x {RAYS} rays, bitmap {RES[0]}x{RES[1]}, {what}"),
        "heliostats_per_gpu": N_HELIOSTATS if scaling == "weak" else N_HELIOSTATS / gpus,
        "surface_points": 4 * POINTS_PER_FACET[0] * POINTS_PER_FACET[1], "rays_per_point": RAYS,
        "bitmap": list(RES), "control_points": list(CONTROL_POINTS),
        "parallelism": (f"heliostat-sharded x{gpus}, NCCL all-reduce of the [T,U,E] flux" if scaling == "weak" else
                        f"one field, samples h -> rank h % {gpus} (HeliostatRayTracer(world_size, rank)), NCCL all-reduce of the "
                        "[T,U,E] flux and of the replicated parameters' gradient; NURBS + alignment replicated per rank as in "
                        "the reference"),
        "l2_policy": "inputs larger than L2 (2.9 GB of distortions/points/normals per step per GPU vs 126 MB L2)",
        "trig": "polynomial sin/cos (<= 1 ulp), FMA-free coordinate path", "accumulate": "fixed-point (deterministic)",
        "alignment": "fused into the trace kernels (orientation applied per point; aligned [N,P,4] tensors never materialised)",
    }
    return cfg


def verify_reduced_flux(wl: "Workload", dev: torch.device, world: int, rank: int, n: int, kind: str, strong: bool) -> dict | None:
    """N > 1, before the timed region, initial parameters: the NCCL-reduced per-target flux and rank 0's parameter gradient
    against a single-process replay on rank 0 (every rank's part traced one after the other, summed in rank order)."""
    wl.opt.zero_grad(set_to_none=True)
    local = wl.forward_local()
    import torch.distributed.nn.functional as dist_fn

    total = dist_fn.all_reduce(local)
    (total * total).mean().backward()
    if strong:
        torch.distributed.all_reduce(wl.param.grad)
    got_total, got_grad = total.detach().clone(), wl.param.grad.detach().clone()
    wl.opt.zero_grad(set_to_none=True)
    out = None
    if rank == 0:
        wl.opt.zero_grad(set_to_none=True)
        mine = wl.forward_local()
        others, grads = [], []
        for r in range(1, world):
            other = Workload(dev, n, world, r, kind=kind, strong=strong)
            part = other.forward_local()
            others.append(part)
            grads.append(other)
        total1 = mine
        for part in others:
            total1 = total1 + (part if strong else part.detach())
        (total1 * total1).mean().backward()
        want_grad = wl.param.grad.detach().clone()
        if strong:      # replicated parameters: sum of the per-rank gradients
            for other in grads:
                want_grad += other.param.grad
        # the autograd-aware all-reduce sums the ranks' (identical) upstream gradients too: every rank's parameter gradient
        # is world x the one-process gradient (the reference divides by the world size after its reduction,
        # artist/optim/surface_reconstructor.py:777)
        want_grad = want_grad * world
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        out = {"flux_max_rel_err": rel(got_total, total1.detach()), "grad_max_rel_err": rel(got_grad, want_grad),
               "what": f"all-reduced [T,U,E] flux over {world} ranks and rank 0's parameter gradient vs a one-process replay of "
                       "every rank's part on rank 0 (same seeds, initial parameters; gradient x world, the all-reduce's autograd "
                       "rule), relative to the peak / largest entry"}
        wl.opt.zero_grad(set_to_none=True)
        del others, grads
        torch.cuda.empty_cache()
    torch.distributed.barrier()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="surface", choices=["surface", "motor"],
                    help="surface = headline (control-point gradients, planar target); motor = BASELINE.json configs[4] "
                         "(motor-position gradients, cylindrical receiver, blocking on)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak = one field per GPU; strong = ONE field split over the ranks by the sampler contract")
    ap.add_argument("--heliostats", type=int, default=None, help="override heliostats per GPU / per field (debug)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-verify", action="store_true", help="N > 1: skip the one-process replay of the reduced flux")
    args = ap.parse_args()
    global N_HELIOSTATS
    if args.heliostats:
        N_HELIOSTATS = args.heliostats
    if args.impl == "reference":
        run_reference(args)
        return

    import __graft_entry__ as entry

    entry.build()
    from artist_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    strong = args.scaling == "strong"
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if distributed:
        torch.distributed.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    warmup = max(3, args.warmup)

    wl = Workload(dev, N_HELIOSTATS, world, rank, kind=args.workload, strong=strong)
    verify = None
    if distributed and not args.skip_verify:
        verify = verify_reduced_flux(wl, dev, world, rank, N_HELIOSTATS, args.workload, strong)
    for _ in range(warmup):
        wl.step()
    # --- device-resident timed region (value) with per-kernel CUDA-event timing and clock sampling ---
    _lib.timing_enabled = True
    _lib.timing_events.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.kernel_launches()
    ms_step = timed_loop(wl.step, args.steps, dev, distributed)
    launches = _lib.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    _lib.timing_enabled = False
    kern_ms = {name: sum(a.elapsed_time(b) for a, b in ev) / len(ev) for name, ev in _lib.timing_events.items()}
    # --- end-to-end leg (host buffers) ---
    for _ in range(2):
        wl.step_e2e()
    wl.flush_e2e()
    ms_e2e = timed_loop(wl.step_e2e, args.steps, dev, distributed, finish=wl.flush_e2e)
    rays = torch.tensor([float(wl.rays_per_step)], device=dev)
    if distributed:
        torch.distributed.all_reduce(rays)

    if rank == 0:
        rays_total = float(rays.item())
        value = rays_total / (ms_step * 1e-3)
        e2e_value = rays_total / (ms_e2e * 1e-3)
        h2d, d2h = wl.e2e_bytes
        peak, peak_src = measured_peak_hbm()
        bf, bb = bytes_per_ray(RAYS, wl.p, RES[0] * RES[1])
        if args.workload == "motor":    # no surface gradients are written: 8 + 32/R + bitmap
            bb = bf
        dom = max(("ab200_trace_fwd", "ab200_trace_bwd"), key=lambda k: kern_ms.get(k, 0.0))
        bpr = bf if dom == "ab200_trace_fwd" else bb
        achieved = wl.rays_per_step * bpr / (kern_ms[dom] * 1e-3) / 1e9
        step_bytes = wl.rays_per_step * (bf + bb)
        out = {
            "metric": "rays/s forward+backward", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world, args.workload, args.scaling),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e,
                    "note": "every step uploads its parameters, incident directions and target indices from pinned host memory "
                            "and downloads its per-target flux, loss and parameter gradient; uploads and downloads run on two "
                            "copy streams and overlap the kernels of the step in flight (double-buffered; the host holds step "
                            "i-1's results when step i has been issued); distortion samples are the tracer's seeded device "
                            "state (as in the reference, which samples them on the device)"},
            "gpu_launches": int(launches),
            "kernel_ms": {k: round(v, 4) for k, v in sorted(kern_ms.items())},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic(dom, wl.rays_per_step) if args.workload == "surface" and not strong else None,
                         "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, "
                                         "profiles/trace_traffic.json)",
                         "algorithmic_bytes_per_launch": wl.rays_per_step * bpr, "peak_source": peak_src,
                         "bytes_per_ray": bpr, "rays_per_launch": wl.rays_per_step,
                         "fwd_bwd_trace_frac": (step_bytes / ((kern_ms["ab200_trace_fwd"] + kern_ms["ab200_trace_bwd"]) * 1e-3) / 1e9) / peak,
                         "whole_step_frac": (step_bytes / (ms_step * 1e-3) / 1e9) / peak},
            "clocks": clocks,
        }
        second = issue_roof(dom, wl.rays_per_step, kern_ms[dom], (clocks or {}).get("sm_mhz")) \
            if args.workload == "surface" and not strong else None   # (the committed capture is of the headline workload)
        if second is not None:
            out["roofline"]["second_roof"] = second
        if verify is not None:
            out["multi_gpu_check"] = verify
        if world == 1 and not args.skip_cpu_baseline:
            from oracle import artist_oracle as O

            threads = os.cpu_count() or 1
            step, rays_cpu = cpu_oracle_step_factory(CPU_SAMPLE_HELIOSTATS, threads, args.workload)
            step()
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                step()
            dt = (time.perf_counter() - t0) / reps
            with O.correctly_rounded_trig():
                step_crt, _ = cpu_oracle_step_factory(CPU_SAMPLE_HELIOSTATS, threads, args.workload)
                step_crt()
            out["parity"] = parity_against_cpu_step(dev, step.first, args.workload, step_crt.first)
            out["cpu_baseline"] = {"value": rays_cpu / dt, "unit": "rays/s", "cores": threads, "kind": "port",
                                   "sample": f"{CPU_SAMPLE_HELIOSTATS} heliostats ({rays_cpu} rays) of the same field per step, "
                                             f"{reps} timed steps after 1 warm-up, torch CPU oracle port with {threads} threads"}
        print(json.dumps(out), flush=True)
    if distributed:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
