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
CPU_SAMPLE_HELIOSTATS = 48   # bounded CPU sample of the same per-heliostat workload
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
    """Device-resident synthetic field + the fwd/bwd step through the public class API."""

    def __init__(self, dev: torch.device, n: int, world: int, rank: int) -> None:
        from artist_b200 import HeliostatRayTracer, NURBSSurfaces, build_synthetic_scenario
        from artist_b200.nurbs import create_nurbs_evaluation_grid

        self.dev, self.n, self.world = dev, n, world
        self.scenario, self.group = build_synthetic_scenario(
            n, number_of_rays=RAYS, points_per_facet=POINTS_PER_FACET, control_points=CONTROL_POINTS,
            surface_bump=SURFACE_BUMP, seed=rank, device=dev)
        g = self.group
        self.mask, self.tidx, self.inc = self.scenario.index_mapping(g)
        self.inc = self.inc.contiguous()
        self.aim = self.scenario.solar_tower.get_centers_of_target_areas(self.tidx)
        self.cp = g.nurbs_control_points.detach().clone().requires_grad_(True)
        g.nurbs_control_points = self.cp
        g.activate_heliostats(self.mask)
        self.surf = NURBSSurfaces(g.nurbs_degrees, self.cp, device=dev)
        grid = create_nurbs_evaluation_grid(torch.tensor(POINTS_PER_FACET), device=dev)
        self.ev = grid[None, None].expand(n, g.number_of_facets_per_heliostat, -1, -1)
        g.align_surfaces_with_incident_ray_directions(self.aim, self.inc, self.mask)
        self.tracer = HeliostatRayTracer(self.scenario, g, blocking_active=False, random_seed=7 + rank,
                                         bitmap_resolution=torch.tensor(RES))
        self.opt = torch.optim.Adam([self.cp], lr=1e-6, fused=True)
        self.p = g.surface_points.shape[1]
        self.rays_per_step = n * self.p * RAYS
        # pinned host mirrors for the end-to-end leg
        self.h_cp = self.cp.detach().cpu().pin_memory()
        self.h_inc = self.inc.cpu().pin_memory()
        self.h_tidx = self.tidx.cpu().pin_memory()
        n_t = int(self.scenario.solar_tower.number_of_target_areas_per_type.sum())
        self.h_flux = torch.empty(n_t, RES[1], RES[0]).pin_memory()
        self.h_grad = torch.empty_like(self.h_cp).pin_memory()
        self.h_loss = torch.empty(1).pin_memory()

    def step(self) -> torch.Tensor:
        g = self.group
        self.opt.zero_grad(set_to_none=True)
        g.activate_heliostats(self.mask)
        pts, nrm = self.surf.calculate_surface_points_and_normals(self.ev, g.active_canting, g.active_facet_translations)
        g.active_surface_points = pts.reshape(self.n, -1, 4)
        g.active_surface_normals = nrm.reshape(self.n, -1, 4)
        g.align_surfaces_with_incident_ray_directions(self.aim, self.inc, self.mask)
        flux, _, _, _ = self.tracer.trace_rays(self.inc, self.mask, self.tidx)
        total = self.tracer.get_bitmaps_per_target(flux, self.tidx)
        if self.world > 1:
            import torch.distributed.nn.functional as dist_fn

            total = dist_fn.all_reduce(total)  # autograd-aware SUM over ranks (NCCL, NVLink)
        loss = (total * total).mean()
        loss.backward()
        self.opt.step()
        self.last_total = total
        return loss

    def step_e2e(self) -> None:
        """Same step with HOST buffers: parameters/inputs come from pinned memory, flux + loss + gradient go back."""
        with torch.no_grad():
            self.cp.copy_(self.h_cp, non_blocking=True)
            self.inc.copy_(self.h_inc, non_blocking=True)
            self.tidx.copy_(self.h_tidx, non_blocking=True)
        loss = self.step_keep_grad()
        self.h_flux.copy_(self.last_total.detach(), non_blocking=True)
        self.h_grad.copy_(self.cp.grad, non_blocking=True)
        self.h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def step_keep_grad(self) -> torch.Tensor:
        loss = self.step()
        return loss

    @property
    def e2e_bytes(self) -> tuple[int, int]:
        h2d = self.h_cp.numel() * 4 + self.h_inc.numel() * 4 + self.h_tidx.numel() * 4
        d2h = self.h_flux.numel() * 4 + self.h_grad.numel() * 4 + 4
        return h2d, d2h


def timed_loop(fn, steps: int, dev: torch.device, distributed: bool) -> float:
    """ms per step over exactly `steps` steps: barrier + synchronize on both sides, CUDA events, max over ranks."""
    if distributed:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if distributed:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.barrier()
        ms = float(t.item())
    return ms / steps


def cpu_oracle_step_factory(n: int, threads: int):
    """The oracle port of the reference's eager path, same per-heliostat workload, on the host cores."""
    from artist_b200.scenario.synthetic import synthetic_field_tensors
    from oracle import artist_oracle as O

    torch.set_num_threads(threads)
    ft = synthetic_field_tensors(n, control_points=CONTROL_POINTS, surface_bump=SURFACE_BUMP, seed=0)
    tg = O.targets_from_field_tensors(ft)
    ev = O.nurbs_evaluation_grid(*POINTS_PER_FACET)[None, None].expand(n, 4, -1, -1)
    tidx = torch.zeros(n, dtype=torch.int32)
    inc = torch.tensor([0.0, 1.0, 0.0, 0.0]).expand(n, -1).contiguous()
    aim = O.aim_points(tg, tidx)
    kin = O.Kin(ft["positions"], ft["translation_deviations"], ft["rotation_deviations"],
                ft["actuator_non_optimizable"], ft["actuator_optimizable"], True)
    cp = ft["nurbs_control_points"].clone().requires_grad_(True)
    opt = torch.optim.Adam([cp], lr=1e-6)
    p = 4 * POINTS_PER_FACET[0] * POINTS_PER_FACET[1]
    du, de = O.sun_distortions(RAYS, p, n, 7)

    first = {}

    def step():
        opt.zero_grad(set_to_none=True)
        pts, nrm = O.nurbs_points_and_normals(cp, 3, 3, ev, ft["canting"], ft["facet_translations"])
        with torch.no_grad():
            ori, _ = O.incident_ray_directions_to_orientations(kin, inc, aim)
        ap, an = O.align_surfaces(pts.reshape(n, -1, 4), nrm.reshape(n, -1, 4), ori)
        flux, *_ = O.trace_rays(ap, an, inc, du, de, tidx, tg, RES, batch_size=min(64, n))
        total = O.bitmaps_per_target(flux, tidx, tg.n_total)
        loss = (total * total).mean()
        loss.backward()
        if not first:   # outputs of the very first step (initial control points): the parity reference of the own arm
            first.update(total=total.detach().clone(), grad=cp.grad.detach().clone(), du=du, de=de)
        opt.step()
        return loss

    step.first = first
    return step, n * p * RAYS


def parity_against_cpu_step(dev: torch.device, first: dict) -> dict:
    """The own arm on the CPU baseline's sample (same field, same distortion samples, first step): per-target flux and
    control-point gradient against the oracle's - the 'flux max rel err' half of BASELINE.json's metric."""
    from artist_b200 import ops

    n = first["du"].shape[0]
    wl = Workload(dev, n, 1, 0)
    ds = wl.tracer.distortions_dataset
    ds.distortions_u, ds.distortions_e = first["du"].to(dev), first["de"].to(dev)
    wl.tracer._packed = ops.pack_distortions(ds.distortions_u, ds.distortions_e)
    wl.step()
    total, grad = wl.last_total.detach().cpu(), wl.cp.grad.detach().cpu()
    return {"flux_max_rel_err": float((total - first["total"]).abs().max() / first["total"].max()),
            "grad_max_rel_err": float((grad - first["grad"]).abs().max() / first["grad"].abs().max()),
            "sample": f"{n} heliostats of the bench field, first step, distortion samples of the CPU baseline step; "
                      "per-target flux relative to its peak, control-point gradient relative to its largest entry; "
                      "CPU side = oracle port (pinned bit-exact to the reference), GPU side = device trig + fixed-point bitmap"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step, rays = cpu_oracle_step_factory(CPU_SAMPLE_HELIOSTATS, threads)
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, note="bounded CPU sample, see cpu_baseline.sample"),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def workload_config(gpus: int, note: str | None = None) -> dict:
    cfg = {
        "workload": (f"synthetic Juelich-scale field, {N_HELIOSTATS} heliostats per GPU x 4 facets x "
                     f"{POINTS_PER_FACET[0]}x{POINTS_PER_FACET[1]} surface points x {RAYS} rays, planar 8x8 m target, "
                     f"bitmap {RES[0]}x{RES[1]}, blocking off; step = NURBS eval + alignment + fused trace + "
                     "per-target flux + loss + backward to control points + Adam"),
        "heliostats_per_gpu": N_HELIOSTATS, "surface_points": 4 * POINTS_PER_FACET[0] * POINTS_PER_FACET[1], "rays_per_point": RAYS,
        "bitmap": list(RES), "control_points": list(CONTROL_POINTS), "parallelism": f"heliostat-sharded x{gpus}, NCCL all-reduce of the [T,U,E] flux",
        "l2_policy": "inputs larger than L2 (2.9 GB of distortions/points/normals per step per GPU vs 126 MB L2)",
        "trig": "polynomial sin/cos (<= 1 ulp), FMA-free coordinate path", "accumulate": "fixed-point (deterministic)",
        "alignment": "fused into the trace kernels (orientation applied per point; aligned [N,P,4] tensors never materialised)",
    }
    if note:
        cfg["note"] = note
    return cfg


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--heliostats", type=int, default=None, help="override heliostats per GPU (debug)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
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
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if distributed:
        torch.distributed.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    warmup = max(3, args.warmup)

    wl = Workload(dev, N_HELIOSTATS, world, rank)
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
    ms_e2e = timed_loop(wl.step_e2e, args.steps, dev, distributed)

    if rank == 0:
        rays_total = wl.rays_per_step * world
        value = rays_total / (ms_step * 1e-3)
        e2e_value = rays_total / (ms_e2e * 1e-3)
        h2d, d2h = wl.e2e_bytes
        peak, peak_src = measured_peak_hbm()
        bf, bb = bytes_per_ray(RAYS, wl.p, RES[0] * RES[1])
        dom = max(("ab200_trace_fwd", "ab200_trace_bwd"), key=lambda k: kern_ms.get(k, 0.0))
        bpr = bf if dom == "ab200_trace_fwd" else bb
        achieved = wl.rays_per_step * bpr / (kern_ms[dom] * 1e-3) / 1e9
        out = {
            "metric": "rays/s forward+backward", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e,
                    "note": "control points, incident directions and target indices uploaded from pinned host memory every "
                            "step; per-target flux, loss and control-point gradient read back; distortion samples are the "
                            "tracer's seeded device state (as in the reference, which samples them on the device)"},
            "gpu_launches": int(launches),
            "kernel_ms": {k: round(v, 4) for k, v in sorted(kern_ms.items())},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(dom, wl.rays_per_step),
                         "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, "
                                         "profiles/trace_traffic.json)",
                         "algorithmic_bytes_per_launch": wl.rays_per_step * bpr, "peak_source": peak_src,
                         "bytes_per_ray": bpr, "rays_per_launch": wl.rays_per_step,
                         "fwd_bwd_trace_frac": (wl.rays_per_step * (bf + bb) / ((kern_ms["ab200_trace_fwd"] + kern_ms["ab200_trace_bwd"]) * 1e-3) / 1e9) / peak},
            "clocks": clocks,
        }
        second = issue_roof(dom, wl.rays_per_step, kern_ms[dom], (clocks or {}).get("sm_mhz"))
        if second is not None:
            out["roofline"]["second_roof"] = second
        if world == 1 and not args.skip_cpu_baseline:
            threads = os.cpu_count() or 1
            step, rays = cpu_oracle_step_factory(CPU_SAMPLE_HELIOSTATS, threads)
            step()
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                step()
            dt = (time.perf_counter() - t0) / reps
            out["parity"] = parity_against_cpu_step(dev, step.first)
            out["cpu_baseline"] = {"value": rays / dt, "unit": "rays/s", "cores": threads, "kind": "port",
                                   "sample": f"{CPU_SAMPLE_HELIOSTATS} heliostats ({rays} rays) of the same field per step, "
                                             f"{reps} timed steps after 1 warm-up, torch CPU oracle port with {threads} threads"}
        print(json.dumps(out), flush=True)
    if distributed:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
