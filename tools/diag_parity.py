"""Diagnostic (GPU box): error tables behind the parity tests - kinematics orientation error vs float64, large-mode
backward gradient error per trig mode, cylinder rows.  Prints plain text."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import artist_oracle as O
from tests import cases
from artist_b200 import ops
from artist_b200.scenario.synthetic import synthetic_field_tensors

dev = torch.device("cuda:0")


def kin_errors(n=300):
    ft = synthetic_field_tensors(n, control_points=(6, 6), surface_bump=0.002)
    g = torch.Generator().manual_seed(5)
    ft["rotation_deviations"] = 0.01 * torch.randn(n, 4, generator=g)
    motor = torch.tensor([[49000.0, 43000.0]]).expand(n, 2) + 3000 * torch.randn(n, 2, generator=g)

    def orc(dtype):
        old = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            cv = lambda x: x.to(dtype) if torch.is_tensor(x) and x.is_floating_point() else x
            f = {k: cv(v) for k, v in ft.items()}
            kin = O.Kin(f["positions"], f["translation_deviations"], f["rotation_deviations"], f["actuator_non_optimizable"],
                        f["actuator_optimizable"], True)
            ang = O.motor_positions_to_angles(kin, cv(motor))
            return O.motor_positions_to_orientations(kin, cv(motor)).double(), ang.double()
        finally:
            torch.set_default_dtype(old)

    o32, a32 = orc(torch.float32)
    o64, a64 = orc(torch.float64)
    from artist_b200.field.kinematics_rigid_body import _initial_orientation_offset

    f = lambda k: ft[k].to(dev)
    og = ops.kinematics_orientations(motor.to(dev), f("rotation_deviations"), f("translation_deviations"), f("actuator_optimizable"),
                                     f("positions"), f("actuator_non_optimizable"), _initial_orientation_offset().to(dev), True)
    og = og.double().cpu()
    rot = lambda o: o[:, :3, :3]
    print("kinematics: rotation-part max abs error vs float64: cpu fp32 %.3e   gpu %.3e   gpu vs cpu32 %.3e" % (
        (rot(o32) - rot(o64)).abs().max(), (rot(og) - rot(o64)).abs().max(), (rot(og) - rot(o32)).abs().max()))
    print("kinematics: translation max abs error vs float64: cpu fp32 %.3e   gpu %.3e" % (
        (o32[:, :3, 3] - o64[:, :3, 3]).abs().max(), (og[:, :3, 3] - o64[:, :3, 3]).abs().max()))
    print("joint angles: cpu fp32 vs float64 %.3e" % (a32 - a64).abs().max())
    worst = (rot(og) - rot(o64)).abs().amax((1, 2)).argmax()
    print("worst heliostat", int(worst), "\n gpu-64:\n", (rot(og) - rot(o64))[worst], "\n cpu32-64:\n", (rot(o32) - rot(o64))[worst])


def bwd_errors():
    from tests.test_gpu_trace_parity import _dev_targets

    n = 300
    case = cases.make_case(n=n, points_per_facet=(12, 12), rays=6)
    res = (256, 256)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, res[1]), torch.linspace(-1, 1, res[0]), indexing="ij")
    wgt = (1.0 + 0.5 * xx - 0.3 * yy + 0.4 * xx * yy + 0.2 * yy * yy)[None].expand(n, -1, -1).contiguous()
    ref, gp, gn = cases.oracle_trace_with_grads(case, res, wgt)
    ref64, gp64, gn64 = cases.oracle_trace_with_grads(case, res, wgt, dtype=torch.float64)
    print("oracle fp32 vs fp64: points %.3e normals %.3e" % ((gp - gp64).abs().max() / gp64.abs().max(),
                                                             (gn - gn64).abs().max() / gn64.abs().max()))
    for nn in (300, 6):
        for trig_mode in (1, 2, 0):
            sl = slice(0, nn)
            opt = ops.TraceOptions(res_e=res[0], res_u=res[1], trig_mode=trig_mode, scatter_sigma=(4.3681e-06) ** 0.5)
            pts = case["points"][sl].to(dev).requires_grad_(True)
            nrm = case["normals"][sl].to(dev).requires_grad_(True)
            trig = cases.cpu_trig(case["dist_u"][sl], case["dist_e"][sl]).to(dev) if trig_mode == 1 else None
            flux, *_ = ops.trace(pts, nrm, case["incident"][sl].to(dev),
                                 ops.pack_distortions(case["dist_u"][sl].to(dev), case["dist_e"][sl].to(dev)),
                                 case["target_idx"][sl].to(dev), _dev_targets(case["targets"], dev), opt, trig=trig)
            (flux * wgt[sl].to(dev)).sum().backward()
            ep = (pts.grad.cpu() - gp[sl]).abs().max() / gp[sl].abs().max()
            en = (nrm.grad.cpu() - gn[sl]).abs().max() / gn[sl].abs().max()
            ep64 = (pts.grad.cpu().double() - gp64[sl]).abs().max() / gp64[sl].abs().max()
            en64 = (nrm.grad.cpu().double() - gn64[sl]).abs().max() / gn64[sl].abs().max()
            fe = (flux.detach().cpu() - ref[sl]).abs().max() / ref[sl].max()
            d = (pts.grad.cpu() - gp[sl]).abs()
            w = d.flatten().argmax()
            print(f"n={nn} trig={trig_mode}: flux {fe:.2e}  grad points vs32 {ep:.2e} vs64 {ep64:.2e}  normals vs32 {en:.2e} vs64 {en64:.2e}"
                  f"  worst at {tuple(int(v) for v in torch.unravel_index(w, d.shape))} gpu {pts.grad.cpu().flatten()[w]:.5e} cpu {gp[sl].flatten()[w]:.5e}")


def motor_chain(n=6):
    """Where does the flux difference of the motor-position chain come from?  Swap the GPU / CPU orientation into the
    other side's trace."""
    from artist_b200.field.kinematics_rigid_body import _initial_orientation_offset
    from tests.test_gpu_trace_parity import _dev_targets

    ppf, rays, res = (12, 12), 6, (128, 128)
    ft = synthetic_field_tensors(n, control_points=(6, 6), surface_bump=0.002)
    g = torch.Generator().manual_seed(5)
    ft["rotation_deviations"] = 0.01 * torch.randn(n, 4, generator=g)
    inc = torch.tensor([0.0, 0.96, -0.28, 0.0]).expand(n, 4).contiguous()
    tidx = torch.zeros(n, dtype=torch.int32)

    def chain(dtype, ori_override=None):
        old = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            cv = lambda x: x.to(dtype) if torch.is_tensor(x) and x.is_floating_point() else x
            f = {k: cv(v) for k, v in ft.items()}
            tg = cases.targets_from(f)
            ev = cv(O.nurbs_evaluation_grid(*ppf))[None, None].expand(n, 4, -1, -1)
            pts, nrm = O.nurbs_points_and_normals(f["nurbs_control_points"], 3, 3, ev, f["canting"], f["facet_translations"])
            kin = O.Kin(f["positions"], f["translation_deviations"], f["rotation_deviations"], f["actuator_non_optimizable"],
                        f["actuator_optimizable"], True)
            if not hasattr(chain, "motor"):
                _, m = O.incident_ray_directions_to_orientations(kin, cv(inc), cases.aim_points(tg, tidx))
                g2 = torch.Generator().manual_seed(9)
                chain.motor = (m.float() + 40.0 * torch.randn(n, 2, generator=g2))
            ori = O.motor_positions_to_orientations(kin, cv(chain.motor)) if ori_override is None else cv(ori_override)
            ap, an = O.align_surfaces(pts.reshape(n, -1, 4), nrm.reshape(n, -1, 4), ori)
            du, de = O.sun_distortions(rays, ap.shape[1], n, 7)
            with O.correctly_rounded_trig():
                ref, *_ = O.trace_rays(ap, an, cv(inc), cv(du), cv(de), tidx, tg, res)
            return ref.sum(0).double(), ori.double(), pts.reshape(n, -1, 4), nrm.reshape(n, -1, 4), tg, du, de
        finally:
            torch.set_default_dtype(old)

    t32, o32, pts32, nrm32, tg32, du, de = chain(torch.float32)
    t64, o64, *_ = chain(torch.float64)
    f = lambda k: ft[k].to(dev)
    og = ops.kinematics_orientations(chain.motor.to(dev), f("rotation_deviations"), f("translation_deviations"),
                                     f("actuator_optimizable"), f("positions"), f("actuator_non_optimizable"),
                                     _initial_orientation_offset().to(dev), True)
    t32_gpu_ori, *_ = chain(torch.float32, ori_override=og.cpu())
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print("motor chain: flux vs float64: cpu32 %.3e | cpu32 trace with GPU orientation %.3e" % (rel(t32, t64), rel(t32_gpu_ori, t64)))
    print("   orientation error vs float64 (rotation part): cpu32 %.3e gpu %.3e ; translation: cpu32 %.3e gpu %.3e" % (
        (o32[:, :3, :3] - o64[:, :3, :3]).abs().max(), (og.double().cpu()[:, :3, :3] - o64[:, :3, :3]).abs().max(),
        (o32[:, :3, 3] - o64[:, :3, 3]).abs().max(), (og.double().cpu()[:, :3, 3] - o64[:, :3, 3]).abs().max()))
    # GPU trace with the CPU32 orientation (fused alignment) and with its own
    opt = ops.TraceOptions(res_e=res[0], res_u=res[1], scatter_sigma=(4.3681e-06) ** 0.5)
    for name, ori in (("cpu32 orientation", o32.float().to(dev)), ("gpu orientation", og)):
        flux, *_ = ops.trace(pts32.to(dev), nrm32.to(dev), inc.to(dev), ops.pack_distortions(du.to(dev), de.to(dev)), tidx.to(dev),
                             _dev_targets(tg32, dev), opt, orientations=ori.contiguous())
        tot = flux.sum(0).double().cpu()
        print("   GPU trace, %s: flux vs float64 %.3e, vs cpu32 %.3e" % (name, rel(tot, t64), rel(tot, t32)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "motor":
        motor_chain()
        sys.exit(0)
    kin_errors()
    bwd_errors()
