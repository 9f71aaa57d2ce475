"""-m gpu: the reference-shaped class API (Scenario / HeliostatGroupRigidBody / HeliostatRayTracer) end to end
against the oracle chain NURBS -> alignment -> trace on the CPU."""
import pytest
import torch

from oracle import artist_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _oracle_chain(ft, ppf, tidx, inc, rays, res, seed=7):
    n = ft["positions"].shape[0]
    tg = cases.targets_from(ft)
    ev = O.nurbs_evaluation_grid(*ppf)[None, None].expand(n, 4, -1, -1)
    pts, nrm = O.nurbs_points_and_normals(ft["nurbs_control_points"], 3, 3, ev, ft["canting"], ft["facet_translations"])
    kin = O.Kin(ft["positions"], ft["translation_deviations"], ft["rotation_deviations"],
                ft["actuator_non_optimizable"], ft["actuator_optimizable"], True)
    ori, motor = O.incident_ray_directions_to_orientations(kin, inc, cases.aim_points(tg, tidx))
    ap, an = O.align_surfaces(pts.reshape(n, -1, 4), nrm.reshape(n, -1, 4), ori)
    return tg, ap, an, motor


def test_flux_prediction_through_class_api():
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario, synthetic_field_tensors

    n, ppf, rays, res = 5, (14, 14), 6, (128, 128)
    ft = synthetic_field_tensors(n, control_points=(8, 8), surface_bump=0.002)
    scenario, group = build_synthetic_scenario(n, number_of_rays=rays, points_per_facet=ppf, device=DEV, field_tensors=ft)
    mask, tidx, inc = scenario.index_mapping(group, single_incident_ray_direction=torch.tensor([0.0, 0.96, -0.28, 0.0]),
                                             single_target_area_index=0)
    group.activate_heliostats(mask)
    group.align_surfaces_with_incident_ray_directions(
        aim_points=scenario.solar_tower.get_centers_of_target_areas(tidx), incident_ray_directions=inc,
        active_heliostats_mask=mask)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor(res))
    flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx)
    per_target = tracer.get_bitmaps_per_target(flux, tidx)
    assert per_target.shape == (2, 128, 128)
    tg, ap, an, motor = _oracle_chain(ft, ppf, tidx.cpu(), inc.cpu(), rays, res)
    assert (group.active_surface_points.cpu() - ap).abs().max() <= 2e-4
    assert ((group.kinematics.active_motor_positions.cpu() - motor).abs() / motor.abs().clamp_min(1)).max() < 2e-4
    du, de = tracer.distortions_dataset.distortions_u.cpu(), tracer.distortions_dataset.distortions_e.cpu()
    # the oracle traces the CUDA-aligned surfaces with the same distortion samples
    ref, ric, rot, _ = O.trace_rays(group.active_surface_points.cpu(), group.active_surface_normals.cpu(), inc.cpu(),
                                    du, de, tidx.cpu(), tg, res)
    assert (flux.cpu() - ref).abs().max() <= 1e-4 * ref.max()
    assert (per_target[0].cpu() - ref.sum(0)).abs().max() <= 1e-4 * ref.sum(0).max()
    assert (ic.cpu() - ric).abs().max() <= 1e-3 and (ot.cpu() - rot).abs().max() <= 1e-3


def test_surface_reconstruction_step_gradients():
    """Config 4 shape: control points -> NURBS -> align -> trace -> loss -> backward to the control points."""
    from artist_b200 import HeliostatRayTracer, NURBSSurfaces, build_synthetic_scenario, synthetic_field_tensors
    from artist_b200.nurbs import create_nurbs_evaluation_grid

    n, ppf, rays, res = 3, (12, 12), 8, (64, 64)
    ft = synthetic_field_tensors(n, control_points=(6, 6), surface_bump=0.002)
    scenario, group = build_synthetic_scenario(n, number_of_rays=rays, points_per_facet=ppf, device=DEV, field_tensors=ft)
    mask, tidx, inc = scenario.index_mapping(group)
    cp = group.nurbs_control_points.detach().clone().requires_grad_(True)
    group.nurbs_control_points = cp
    group.activate_heliostats(mask)
    surf = NURBSSurfaces(group.nurbs_degrees, group.active_nurbs_control_points, device=torch.device(DEV))
    ev = create_nurbs_evaluation_grid(torch.tensor(ppf), device=torch.device(DEV))[None, None].expand(n, 4, -1, -1)
    pts, nrm = surf.calculate_surface_points_and_normals(ev, group.active_canting, group.active_facet_translations)
    group.active_surface_points, group.active_surface_normals = pts.reshape(n, -1, 4), nrm.reshape(n, -1, 4)
    group.align_surfaces_with_incident_ray_directions(scenario.solar_tower.get_centers_of_target_areas(tidx), inc, mask)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor(res))
    flux, *_ = tracer.trace_rays(inc, mask, tidx)
    # A SMOOTH functional of the flux: the orientation of the device alignment differs from the CPU one in the last bit
    # for some heliostats (transcendental functions), a handful of rays then land in the neighbouring pixel, and with
    # independent random weights per pixel ONE such flip moves a control point's gradient by several per cent
    # (measured: 4.7e-2 with torch.rand weights vs 2.2e-4 with these, same kernels, tools/diag note in DESIGN §2) - the
    # trace-level tests compare gradients under random weights on identical aligned inputs instead.
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, res[1]), torch.linspace(-1, 1, res[0]), indexing="ij")
    wgt = (torch.exp(-((xx - 0.2) ** 2 + (yy + 0.1) ** 2) / 0.3) + 0.3 * xx)[None].expand(n, -1, -1).contiguous()
    (flux * wgt.to(DEV)).sum().backward()
    # oracle: same chain with autograd on the CPU, same distortions
    cpo = ft["nurbs_control_points"].clone().requires_grad_(True)
    tg = cases.targets_from(ft)
    evo = O.nurbs_evaluation_grid(*ppf)[None, None].expand(n, 4, -1, -1)
    po, no = O.nurbs_points_and_normals(cpo, 3, 3, evo, ft["canting"], ft["facet_translations"])
    kin = O.Kin(ft["positions"], ft["translation_deviations"], ft["rotation_deviations"],
                ft["actuator_non_optimizable"], ft["actuator_optimizable"], True)
    ori, _ = O.incident_ray_directions_to_orientations(kin, inc.cpu(), cases.aim_points(tg, tidx.cpu()))
    ap, an = O.align_surfaces(po.reshape(n, -1, 4), no.reshape(n, -1, 4), ori)
    ref, *_ = O.trace_rays(ap, an, inc.cpu(), tracer.distortions_dataset.distortions_u.cpu(),
                           tracer.distortions_dataset.distortions_e.cpu(), tidx.cpu(), tg, res)
    (ref * wgt).sum().backward()
    scale = cpo.grad.abs().max()
    err = (cp.grad.cpu() - cpo.grad).abs().max() / scale
    assert err <= 1e-3, f"control-point gradient error {err:.3e}"     # measured 2.2e-4 (round 1 accepted 5e-3)


def test_motor_position_gradients_and_mask_assertion():
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario

    scenario, group = build_synthetic_scenario(4, number_of_rays=5, points_per_facet=(10, 10), device=DEV)
    mask, tidx, inc = scenario.index_mapping(group)
    group.activate_heliostats(mask)
    group.align_surfaces_with_incident_ray_directions(scenario.solar_tower.get_centers_of_target_areas(tidx), inc, mask)
    motor = group.kinematics.active_motor_positions.detach().clone().requires_grad_(True)
    group.activate_heliostats(mask)
    group.align_surfaces_with_motor_positions(motor, mask)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor([64, 64]))
    flux, *_ = tracer.trace_rays(inc, mask, tidx)
    col = torch.linspace(0, 1, 64, device=DEV)
    (flux * col).sum().backward()
    assert torch.isfinite(motor.grad).all() and motor.grad.abs().max() > 0
    with pytest.raises(AssertionError, match="not aligned"):
        tracer.trace_rays(inc, torch.zeros_like(mask), tidx)
    blocked = HeliostatRayTracer(scenario, group, blocking_active=True, bitmap_resolution=torch.tensor([64, 64]))
    fb, _, _, bl = blocked.trace_rays(inc, mask, tidx)      # default blocking_active=True works (no shadowing here)
    assert torch.equal(fb, flux) and (bl == 1).all()


@pytest.mark.parametrize("n", [6, 300])
def test_motor_position_and_deviation_gradients_vs_oracle_autograd(n):
    """Config 5 shape (``artist/optim/aim_point_optimizer.py:384-402``): motor positions -> ``align_surfaces_with_motor_
    positions`` (kinematics kernel) -> fused trace -> loss; gradients w.r.t. the motor positions and the kinematic
    rotation deviations against the oracle's autograd through the same chain.  n = 300 runs the one-CTA-per-sample trace
    kernels (dL/dO reduced in-kernel), n = 6 the split-mode ones (dL/dO by atomics)."""
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario, synthetic_field_tensors

    ppf, rays, res = (12, 12), 6, (128, 128)
    ft = synthetic_field_tensors(n, control_points=(6, 6), surface_bump=0.002)
    g = torch.Generator().manual_seed(5)
    ft["rotation_deviations"] = 0.01 * torch.randn(n, 4, generator=g)
    scenario, group = build_synthetic_scenario(n, number_of_rays=rays, points_per_facet=ppf, device=DEV, field_tensors=ft)
    mask, tidx, inc = scenario.index_mapping(group, single_incident_ray_direction=torch.tensor([0.0, 0.96, -0.28, 0.0]),
                                             single_target_area_index=0)
    group.activate_heliostats(mask)
    group.align_surfaces_with_incident_ray_directions(scenario.solar_tower.get_centers_of_target_areas(tidx), inc, mask)
    motor0 = group.kinematics.active_motor_positions.detach().clone()
    # de-tune the motors a little so that the spots sit off-centre (non-trivial gradients)
    g2 = torch.Generator().manual_seed(9)
    motor0 = motor0 + (40.0 * torch.randn(n, 2, generator=g2)).to(DEV)
    motor = motor0.clone().requires_grad_(True)
    rot = group.kinematics.rotation_deviation_parameters.detach().clone().requires_grad_(True)
    group.kinematics.rotation_deviation_parameters = rot
    group.activate_heliostats(mask)
    group.align_surfaces_with_motor_positions(motor, mask)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor(res))
    flux, *_ = tracer.trace_rays(inc, mask, tidx)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, res[1]), torch.linspace(-1, 1, res[0]), indexing="ij")
    wgt = 1.0 + 0.6 * xx - 0.4 * yy + 0.3 * xx * yy
    total = tracer.get_bitmaps_per_target(flux, tidx)[0]
    (total * wgt.to(DEV)).sum().backward()
    assert motor.grad is not None and rot.grad is not None
    # ---- oracle: identical chain on the CPU with autograd, same distortion samples; float32 (= the reference's arithmetic)
    # and float64 (conditioning-free gold).  The chain motor position -> law of cosines -> nine 4x4 products -> 60 m lever arm
    # loses ~3 digits in fp32 (a last-bit change of an actuator angle moves the spot by ~1e-3 pixel), so two CORRECT fp32
    # implementations differ from each other by about as much as each differs from the exact result: the kernel is required
    # to be no further from the float64 result than twice the reference arithmetic's own distance (+ the fp32 trace floor).
    def oracle(dtype):
        old = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            cv = lambda x: x.to(dtype) if torch.is_tensor(x) and x.is_floating_point() else x
            f = {k: cv(v) for k, v in ft.items()}
            tg = cases.targets_from(f)
            ev = cv(O.nurbs_evaluation_grid(*ppf))[None, None].expand(n, 4, -1, -1)
            pts, nrm = O.nurbs_points_and_normals(f["nurbs_control_points"], 3, 3, ev, f["canting"], f["facet_translations"])
            mo = cv(motor0.cpu()).clone().requires_grad_(True)
            ro = f["rotation_deviations"].clone().requires_grad_(True)
            kin = O.Kin(f["positions"], f["translation_deviations"], ro, f["actuator_non_optimizable"],
                        f["actuator_optimizable"], True)
            ori = O.motor_positions_to_orientations(kin, mo)
            ap, an = O.align_surfaces(pts.reshape(n, -1, 4), nrm.reshape(n, -1, 4), ori)
            ref, *_ = O.trace_rays(ap, an, cv(inc.cpu()), cv(tracer.distortions_dataset.distortions_u.cpu()),
                                   cv(tracer.distortions_dataset.distortions_e.cpu()), tidx.cpu(), tg, res)
            tot = ref.sum(0)
            (tot * cv(wgt)).sum().backward()
            return tot.detach().double(), mo.grad.double(), ro.grad.double()
        finally:
            torch.set_default_dtype(old)

    t32, gm32, gr32 = oracle(torch.float32)     # the reference's arithmetic (the device polynomial reproduces torch's cos)
    t64, gm64, gr64 = oracle(torch.float64)
    rel = lambda x, gold: float((x.double().cpu() - gold).abs().max() / gold.abs().max())
    own = (rel(total.detach(), t64), rel(motor.grad, gm64), rel(rot.grad, gr64))
    ref32 = (rel(t32, t64), rel(gm32, gm64), rel(gr32, gr64))
    # (two fp32 evaluations of this chain scatter around the exact result: measured with tools/diag_parity.py, the CPU
    # oracle's and the kernels' distances to float64 are 1e-4 ... 8e-4 of the peak pixel depending on the heliostats; the
    # trace itself, fed the oracle's orientations, reproduces the oracle to 2e-6)
    caps = (1.5e-3, 5e-4, 5e-4)
    for name, o, r32, cap in zip(("flux", "motor-position gradient", "rotation-deviation gradient"), own, ref32, caps):
        assert o <= max(10.0 * r32, 2e-4) and o <= cap, f"{name}: error vs float64 {o:.2e}, the fp32 oracle's own {r32:.2e}"
    # and, in absolute terms, close to the fp32 oracle as well
    # (measured, the two parametrisations: flux 8e-5 / 7.8e-4 of peak, gradients 2.5e-6 / 1.7e-4 - the second one has a
    # heliostat whose orientation differs from torch-CPU's in the last bit; round 1 bars: 3e-3 and 5e-3)
    assert rel(total.detach(), t32.double()) <= 1.5e-3 and rel(motor.grad, gm32) <= 5e-4 and rel(rot.grad, gr32) <= 5e-4


def test_lazy_alignment_fused_and_materialised_paths_agree():
    """``align_surfaces_with_*`` only records the orientation; the tracer fuses the rotation into its kernels.
    Reading ``active_surface_points`` materialises the aligned tensors: same flux bit for bit, and the aligned
    tensors equal an explicit ``align_surfaces`` call."""
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario, ops

    scenario, group = build_synthetic_scenario(4, number_of_rays=5, points_per_facet=(10, 10), device=DEV)
    mask, tidx, inc = scenario.index_mapping(group)
    aim = scenario.solar_tower.get_centers_of_target_areas(tidx)
    group.activate_heliostats(mask)
    raw_p, raw_n = group.active_surface_points, group.active_surface_normals
    group.align_surfaces_with_incident_ray_directions(aim, inc, mask)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor([64, 64]))
    pending = group._fused_alignment()
    assert pending is not None and pending[0] is raw_p                     # nothing aligned in memory yet
    fused = tracer.trace_rays(inc, mask, tidx)
    assert group._fused_alignment() is not None
    ap, an = ops.align_surfaces(raw_p, raw_n, pending[2])
    assert torch.equal(group.active_surface_points, ap) and torch.equal(group.active_surface_normals, an)
    assert group._fused_alignment() is None                                # materialised by the reads above
    eager = tracer.trace_rays(inc, mask, tidx)
    for a, b in zip(fused, eager):
        assert torch.equal(a, b)
    refl = group.preferred_reflection_directions
    assert refl.shape == an.shape and torch.isfinite(refl).all()
    # a fresh activation drops the recorded alignment without running it
    group.activate_heliostats(mask)
    assert group._fused_alignment() is None and group.active_surface_points is raw_p


@pytest.mark.parametrize("blocking,one_cta", [(False, False), (False, True), (True, False), (True, True)])
def test_activation_index_map_equals_replicated_copies(blocking, one_cta):
    """``activate_heliostats`` with a selecting / replicating mask (``artist/field/heliostat_group.py:256-315``): the tracer
    reads the group's un-replicated surfaces through an index map (``ab200_trace_args::src_rows``) instead of
    ``repeat_interleave`` copies.  Flux and factors are bit-identical to tracing the materialised copies; the gradient w.r.t.
    the group's surface points (``ab200_replica_sum`` of the per-sample rows) equals autograd's ``index_select`` backward;
    without surface gradients (motor-position optimisation) the backward skips them and returns the same orientation
    gradient."""
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario, synthetic_field_tensors

    n, ppf, rays, res = 6, (10, 10), 6, (64, 64)
    ft = synthetic_field_tensors(n, control_points=(6, 6), surface_bump=0.002)
    scenario, group = build_synthetic_scenario(n, number_of_rays=rays, points_per_facet=ppf, device=DEV, field_tensors=ft)
    mask = torch.tensor([2, 0, 1, 3, 0, 1], dtype=torch.int32, device=DEV)
    n_s = int(mask.sum())
    tidx = torch.zeros(n_s, dtype=torch.int32, device=DEV)
    inc = torch.nn.functional.normalize(torch.tensor([[0.0, 0.96, -0.28, 0.0]], device=DEV)
                                        + 0.05 * torch.randn(n_s, 4, device=DEV) * torch.tensor([1.0, 1.0, 1.0, 0.0], device=DEV), dim=1)
    aim = scenario.solar_tower.get_centers_of_target_areas(tidx)
    wgt = torch.rand(n_s, res[1], res[0], device=DEV)
    base_p, base_n = group.surface_points, group.surface_normals

    def run(materialise: bool, surface_grads: bool = True):
        group.surface_points = base_p.detach().clone().requires_grad_(surface_grads)
        group.surface_normals = base_n.detach().clone().requires_grad_(surface_grads)
        group.activate_heliostats(mask)
        if materialise:
            assert group.active_surface_points.shape[0] == n_s      # the read gathers the copies (reference behaviour)
            assert group.active_surface_normals.shape[0] == n_s
        else:
            assert group._pending_gather is not None and group._asp is None
        motors = aligned_motors.clone().requires_grad_(True)
        group.align_surfaces_with_motor_positions(motors + 0.0, mask)
        tracer = HeliostatRayTracer(scenario, group, blocking_active=blocking, bitmap_resolution=torch.tensor(res))
        tracer._force_one_cta_per_sample = one_cta      # (the kernels a full field runs: MAP instantiations of both modes)
        if not materialise:
            assert group._fused_alignment(with_map=True)[3] is not None, "the index map was meant to reach the tracer"
        out = tracer.trace_rays(inc, mask, tidx)
        (out[0] * wgt).sum().backward()
        return out, group.surface_points.grad, group.surface_normals.grad, motors.grad

    group.activate_heliostats(mask)
    group.align_surfaces_with_incident_ray_directions(aim, inc, mask)   # sets the motor positions of the samples
    aligned_motors = group.kinematics.active_motor_positions.detach().clone()
    lazy, gp, gn, gm = run(False)
    eager, gp_e, gn_e, gm_e = run(True)
    assert lazy[0].sum() > 0
    for a, b in zip(lazy, eager):
        assert torch.equal(a, b)
    assert gp[1].abs().max() == 0 and gp[4].abs().max() == 0            # de-activated heliostats
    for a, b in ((gp, gp_e), (gn, gn_e), (gm, gm_e)):
        assert (a - b).abs().max() <= 1e-5 * b.abs().max()
    _, none_p, none_n, gm_only = run(False, surface_grads=False)
    assert none_p is None and none_n is None
    if blocking:    # the blockers' corner rows feed the motor gradient too: equal up to their summation order
        assert (gm_only - gm).abs().max() <= 1e-5 * gm.abs().max()
    else:
        assert torch.equal(gm_only, gm)
    group.surface_points, group.surface_normals = base_p, base_n
