"""Evidence run: the UPSTREAM package's own callers, unmodified, on top of artist_b200 (``artist_b200.compat``).

Needs a GPU and the upstream package importable (``ARTIST_REFERENCE_ROOT``, ``/root/reference`` or ``baseline/_ref``);
it is a one-off evidence tool - nothing in ``tests/``, ``bench.py`` or ``smoke()`` imports it.  What it does:

 1. imports the upstream ``artist`` package (h5py -> ``artist_b200.io.h5lite``, colorlog / paint stubs);
 2. runs upstream's ``AimPointOptimizer`` (``artist/optim/aim_point_optimizer.py``, the tutorial-05 caller:
    motor-position optimisation through ``HeliostatRayTracer`` with blocking, KL loss, Adam + scheduler + early stopping)
    on the reference's own scenario file ``test_scenario_paint_four_heliostats.h5`` TWICE on the same GPU:
      a) as shipped - upstream classes, PyTorch CUDA-eager kernels;
      b) after ``compat.install()`` - the very same upstream optimiser code, artist_b200 classes underneath;
 3. does the same for the tutorial-01 / 02 flow (activate, align, trace, per-target bitmaps);
 4. prints the loss trajectories, the final flux and motor positions of both runs, their differences and wall times.

usage:  python tools/run_upstream_callers.py [epochs]
"""
from __future__ import annotations

import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


TARGET = int(os.environ.get("AB200_UPSTREAM_TARGET", "3"))   # 3 = the cylindrical receiver (tutorial 05); 0..2 planar


def _import_upstream():
    for cand in (os.environ.get("ARTIST_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "artist")):
            os.environ["ARTIST_REFERENCE_ROOT"] = cand
            break
    else:
        raise SystemExit("upstream package not found (set ARTIST_REFERENCE_ROOT)")
    from artist_b200.io import h5lite

    h5 = types.ModuleType("h5py")
    h5.File = h5lite.File
    h5.Group = h5lite.Group
    h5.Dataset = getattr(h5lite, "Dataset", object)
    sys.modules["h5py"] = h5
    import importlib

    ref_import = importlib.import_module("tools.ref_import")
    ref_import.REFERENCE_ROOT = os.environ["ARTIST_REFERENCE_ROOT"]
    return ref_import.import_reference()


def _sync(device) -> None:
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()


def _scenario_file() -> str:
    golden = torch.load(os.path.join(ROOT, "tests", "golden", "scenario_golden.pt"), weights_only=False)
    path = "/tmp/upstream_four_heliostats.h5"
    with open(path, "wb") as fh:
        fh.write(golden["four_heliostats"]["file_bytes"].numpy().tobytes())
    return path


def _ddp_setup(device):
    return {"device": device, "is_distributed": False, "is_nested": False, "rank": 0, "world_size": 1, "process_subgroup": None,
            "groups_to_ranks_mapping": {0: [0, 1]}, "heliostat_group_rank": 0, "heliostat_group_world_size": 1,
            "ranks_to_groups_mapping": {0: [0], 1: [0]}}


def run_aim_point_optimizer(path: str, device, epochs: int):
    import h5py
    from artist.flux import bitmap
    from artist.optim.aim_point_optimizer import AimPointOptimizer
    from artist.optim.loss import KLDivergenceLoss
    from artist.scenario.scenario import Scenario
    from artist.util import constants, indices

    with h5py.File(path, "r") as fh:
        scenario = Scenario.load_scenario_from_hdf5(scenario_file=fh, device=device)
    scenario.set_number_of_rays(number_of_rays=4)
    res = torch.tensor([256, 256], device=device)
    cfg = {
        constants.optimization: {constants.initial_learning_rate: 3e-4, constants.tolerance: 0.0005, constants.max_epoch: epochs,
                                 constants.batch_size: 50, constants.log_step: 1, constants.early_stopping_delta: 1e-4,
                                 constants.early_stopping_patience: 100, constants.early_stopping_window: 100},
        constants.scheduler: {constants.scheduler_type: constants.reduce_on_plateau, constants.gamma: 0.9, constants.lr_min: 1e-6,
                              constants.lr_max: 1e-3, constants.step_size_up: 500, constants.reduce_factor: 0.3,
                              constants.patience: 100, constants.threshold: 1e-3, constants.cooldown: 10},
        constants.constraints: {constants.rho_flux_integral: 1.0, constants.rho_local_flux: 1.0, constants.rho_intercept: 1.0,
                                constants.max_flux_density: 1000000},
    }
    dni = 800
    canting_norm = (torch.norm(scenario.heliostat_field.heliostat_groups[0].canting[0], dim=1)[0])[:2]
    area = ((canting_norm * 4) + 0.02).prod() * scenario.heliostat_field.number_of_heliostats_per_group.sum()
    e_t = bitmap.trapezoid_distribution(total_width=res[indices.unbatched_bitmap_e], slope_width=30, plateau_width=110, device=device)
    u_t = bitmap.trapezoid_distribution(total_width=res[indices.unbatched_bitmap_u], slope_width=30, plateau_width=110, device=device)
    gt = u_t.unsqueeze(indices.unbatched_bitmap_u) * e_t.unsqueeze(indices.unbatched_bitmap_e)
    gt = gt / gt.sum() * (dni * area * 0.75)
    if os.environ.get("AB200_UPSTREAM_BLOCKING", "1") == "0":   # diagnosis only: the optimiser's tracer without blocking
        import functools

        import artist.optim.aim_point_optimizer as apo

        tracer_cls = apo.HeliostatRayTracer.func if isinstance(apo.HeliostatRayTracer, functools.partial) else apo.HeliostatRayTracer
        apo.HeliostatRayTracer = functools.partial(tracer_cls, blocking_active=False)
    opt = AimPointOptimizer(ddp_setup=_ddp_setup(device), scenario=scenario, optimization_configuration=cfg,
                            incident_ray_direction=torch.tensor([0.0, 1.0, 0.0, 0.0], device=device), target_area_index=TARGET,
                            ground_truth=gt, dni=dni, bitmap_resolution=res, device=device)
    # epoch-0 gradient of the flux loss w.r.t. the optimiser's own parameters, through the optimiser's own methods
    (params, scales, initial, masks, tidxs, incs, offsets) = opt._initialize_group_parameters(device=device)
    probe = torch.optim.Adam(params, lr=1e-3)
    opt._align_all_groups(optimizer=probe, initial_motor_positions_all_groups=initial, scales_all_groups=scales,
                          active_heliostats_masks_all_groups=masks, device=device)
    total_flux, *_ = opt._trace_and_accumulate_flux(active_heliostats_masks_all_groups=masks, target_area_indices_all_groups=tidxs,
                                                    incident_ray_directions_all_groups=incs, group_offsets=offsets, device=device)
    flux_loss = KLDivergenceLoss()(prediction=total_flux.unsqueeze(indices.heliostat_dimension),
                                   ground_truth=gt.unsqueeze(indices.heliostat_dimension),
                                   target_area_indices=torch.tensor([TARGET], device=device),
                                   reduction_dimensions=(indices.batched_bitmap_e, indices.batched_bitmap_u), device=device)
    flux_loss.sum().backward(retain_graph=True)
    g_kl = [p.grad.detach().float().cpu().clone() for p in params]
    for p in params:
        p.grad = None
    # a well-conditioned functional of the same flux (the KL loss divides by prediction + 1e-12: pixels at the fringe of the
    # focal spot dominate its gradient, which makes it sensitive to single rays)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 256, device=device), torch.linspace(-1, 1, 256, device=device), indexing="ij")
    (total_flux * (1.0 + 0.5 * xx - 0.3 * yy + 0.4 * xx * yy + 0.2 * yy * yy)).sum().backward()
    g_smooth = [p.grad.detach().float().cpu().clone() for p in params]
    run_aim_point_optimizer.epoch0 = (total_flux.detach().float().cpu(), g_kl, g_smooth)
    _sync(device)
    t0 = time.perf_counter()
    result = opt.optimize(loss_definition=KLDivergenceLoss(), device=device)
    _sync(device)
    dt = time.perf_counter() - t0
    motors = [g.kinematics.motor_positions.detach().float().cpu().clone() for g in scenario.heliostat_field.heliostat_groups]
    classes = sorted({type(g).__module__ + "." + type(g).__name__ for g in scenario.heliostat_field.heliostat_groups})
    return result, motors, dt, classes


def run_flux_prediction(path: str, device):
    import h5py
    from artist.raytracing.heliostat_ray_tracer import HeliostatRayTracer
    from artist.scenario.scenario import Scenario

    with h5py.File(path, "r") as fh:
        scenario = Scenario.load_scenario_from_hdf5(scenario_file=fh, device=device)
    total = None
    t0 = time.perf_counter()
    for group in scenario.heliostat_field.heliostat_groups:
        mask, tidx, inc = scenario.index_mapping(heliostat_group=group, single_incident_ray_direction=torch.tensor(
            [0.0, 1.0, 0.0, 0.0], device=device), single_target_area_index=1, device=device)
        group.activate_heliostats(active_heliostats_mask=mask, device=device)
        group.align_surfaces_with_incident_ray_directions(
            aim_points=scenario.solar_tower.get_centers_of_target_areas(target_area_indices=tidx, device=device),
            incident_ray_directions=inc, active_heliostats_mask=mask, device=device)
        tracer = HeliostatRayTracer(scenario=scenario, heliostat_group=group, batch_size=10,
                                    bitmap_resolution=torch.tensor([256, 256], device=device))
        flux, _, _, _ = tracer.trace_rays(incident_ray_directions=inc, active_heliostats_mask=mask,
                                          target_area_indices=tidx, device=device)
        per_target = tracer.get_bitmaps_per_target(bitmaps_per_heliostat=flux, target_area_indices=tidx, device=device)
        total = per_target if total is None else total + per_target
    _sync(device)
    return total.detach().float().cpu(), time.perf_counter() - t0, type(tracer).__module__


def main() -> None:
    epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    device = torch.device("cuda:0")
    _import_upstream()
    import artist.optim.aim_point_optimizer  # noqa: F401  (loads every upstream module the caller needs)

    import __graft_entry__ as entry

    entry.build()
    from artist_b200 import compat

    path = _scenario_file()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())

    flux_up, t_up, mod_up = run_flux_prediction(path, device)
    res_up, motors_up, dt_up, cls_up = run_aim_point_optimizer(path, device, epochs)
    e0_up = run_aim_point_optimizer.epoch0
    res_up2, _, _, _ = run_aim_point_optimizer(path, device, epochs)
    e0_up2 = run_aim_point_optimizer.epoch0
    res_cpu, _, dt_cpu, _ = run_aim_point_optimizer(path, torch.device("cpu"), epochs)     # the reference on its other device (CPU RNG: other distortion samples)
    e0_cpu = run_aim_point_optimizer.epoch0
    changed = compat.install()
    print(f"compat.install(): {len(changed)} bindings rebound, e.g. {changed[:6]}")
    flux_own, t_own, mod_own = run_flux_prediction(path, device)
    flux_own2, t_own2, _ = run_flux_prediction(path, device)
    res_own, motors_own, dt_own, cls_own = run_aim_point_optimizer(path, device, epochs)
    e0_own = run_aim_point_optimizer.epoch0
    res_own2, _, dt_own2, _ = run_aim_point_optimizer(path, device, epochs)
    e0_own2 = run_aim_point_optimizer.epoch0
    compat.uninstall()

    print(f"[tutorial 01/02 flow] tracer module upstream run: {mod_up}; after install: {mod_own}")
    print(f"  per-target flux, max |diff| / peak: {rel(flux_own, flux_up):.3e}  (sum upstream {float(flux_up.sum()):.4f}, "
          f"artist_b200 {float(flux_own.sum()):.4f});  wall {t_up * 1e3:.1f} ms upstream CUDA-eager vs {t_own2 * 1e3:.1f} ms "
          f"(first call incl. lazy init {t_own * 1e3:.1f} ms)")
    print(f"[AimPointOptimizer.optimize, {epochs} epochs, target area {TARGET}] group classes upstream run: {cls_up}; after install: {cls_own}")

    def flat(x):
        if torch.is_tensor(x):
            return x.detach().float().cpu().reshape(-1)
        return torch.tensor([float(x)])

    names = ("final loss", "result[1]", "result[2]", "result[3]", "result[4]")
    for name, a, b in zip(names, res_up, res_own):
        try:
            fa, fb = flat(a), flat(b)
            if fa.numel() == fb.numel() and fa.numel() > 0:
                print(f"  {name}: upstream {fa[:4].tolist()} artist_b200 {fb[:4].tolist()}  max rel diff {rel(fb, fa):.3e}")
        except Exception as exc:   # result entries that are not numeric
            print(f"  {name}: not compared ({type(a).__name__}: {exc})")
    print(f"  epoch-0 total flux (sum over groups, blocking on): max |diff| / peak {rel(e0_own[0], e0_up[0]):.3e}")
    for i, (gu, go, gc) in enumerate(zip(e0_up[2], e0_own[2], e0_cpu[2])):
        print(f"  epoch-0 d(smooth functional of the flux)/d(reparameterised motor positions), group {i}: artist_b200 vs upstream-CUDA "
              f"max |diff| / max |grad| {rel(go, gu):.3e}; upstream-CPU (other distortion samples: mt19937, not Philox) vs upstream-CUDA {rel(gc, gu):.3e}")
    for i, (gu, go, gu2, go2, gc) in enumerate(zip(e0_up[1], e0_own[1], e0_up2[1], e0_own2[1], e0_cpu[1])):
        print(f"  epoch-0 d(KL flux loss)/d(reparameterised motor positions), group {i}: artist_b200 vs upstream-CUDA {rel(go, gu):.3e}; "
              f"upstream-CPU (other distortion samples: mt19937, not Philox) vs upstream-CUDA {rel(gc, gu):.3e} (run-to-run: upstream {rel(gu2, gu):.1e}, artist_b200 {rel(go2, go):.1e})")
    print(f"  epoch-0 total flux: upstream-CPU (other distortion samples: mt19937, not Philox) vs upstream-CUDA max |diff| / peak {rel(e0_cpu[0], e0_up[0]):.3e}")
    hu, ho, ho2, hu2 = res_up[1], res_own[1], res_own2[1], res_up2[1]
    for key in ("total_loss", "flux_loss", "flux_integral"):
        if key in hu and key in ho:
            fmt = lambda xs: "[" + ", ".join(f"{x:.6g}" for x in xs) + "]"
            print(f"  {key} per epoch (first 4 / last 2): upstream {fmt(hu[key][:4])} ... {fmt(hu[key][-2:])}")
            print(f"  {' ' * len(key)}                          artist_b200 {fmt(ho[key][:4])} ... {fmt(ho[key][-2:])}")
            n = min(len(hu[key]), len(ho[key]))
            d = [abs(a - b) / max(abs(a), 1e-30) for a, b in zip(hu[key][:n], ho[key][:n])]
            print(f"  {' ' * len(key)}   rel diff per epoch: first {d[0]:.2e}, epoch 5 {d[min(5, n - 1)]:.2e}, last {d[-1]:.2e}; "
                  f"run-to-run identical: upstream {hu[key] == hu2[key]}, artist_b200 {ho[key] == ho2[key]}")
            hc = res_cpu[1]
            d2 = [abs(a - b) / max(abs(a), 1e-30) for a, b in zip(hu[key], hc[key])]
            print(f"  {' ' * len(key)}   upstream on the CPU (other distortion samples) vs upstream on CUDA: first {d2[0]:.2e}, epoch 5 {d2[min(5, len(d2) - 1)]:.2e}, "
                  f"last {d2[-1]:.2e}   (upstream CPU {fmt(hc[key][:4])} ... {fmt(hc[key][-2:])})")
    for i, (a, b) in enumerate(zip(motors_up, motors_own)):
        print(f"  optimised motor positions, group {i}: max |diff| {float((a - b).abs().max()):.3f} increments of "
              f"{float(a.abs().max()):.0f} (rel {rel(b, a):.3e})")
    print(f"  wall time for {epochs} epochs: upstream on the CPU ({torch.get_num_threads()} threads) {dt_cpu:.2f} s; upstream CUDA-eager {dt_up:.2f} s; artist_b200 underneath {dt_own2:.2f} s "
          f"(first run {dt_own:.2f} s) -> {dt_up / dt_own2:.1f}x")


if __name__ == "__main__":
    main()
