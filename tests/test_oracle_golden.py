"""CPU (-m "not gpu"): the oracle against fixtures produced by the REAL reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import pytest
import torch

from oracle import artist_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_golden.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def _targets(field):
    return O.targets_from_field_tensors(field)


@pytest.mark.parametrize("key", ["trace_linear", "trace_ideal"])
def test_trace_pipeline_bit_exact(golden, key):
    g = golden[key]
    f = g["field"]
    tg = _targets(f)
    n = f["positions"].shape[0]
    linear = bool((f["actuator_non_optimizable"][:, 0] == 0).all())
    ev = O.nurbs_evaluation_grid(*g["points_per_facet"])[None, None].expand(n, 4, -1, -1)
    pts, nrm = O.nurbs_points_and_normals(f["nurbs_control_points"], 3, 3, ev, f["canting"], f["facet_translations"])
    assert torch.equal(pts.reshape(n, -1, 4), g["surface_points"])
    assert torch.equal(nrm.reshape(n, -1, 4), g["surface_normals"])
    kin = O.Kin(f["positions"], f["translation_deviations"], f["rotation_deviations"], f["actuator_non_optimizable"],
                f["actuator_optimizable"], linear)
    assert torch.equal(O.aim_points(tg, g["target_idx"]), g["aim"])
    ori, motor = O.incident_ray_directions_to_orientations(kin, g["incident"], g["aim"])
    assert torch.equal(motor, g["motor"]) and torch.equal(ori, g["orientations_incident"])
    assert torch.equal(O.motor_positions_to_orientations(kin, g["motor"]), g["orientations_motor"])
    ap, an = O.align_surfaces(g["surface_points"], g["surface_normals"], ori)
    assert torch.equal(ap, g["aligned_points"]) and torch.equal(an, g["aligned_normals"])
    du, de = O.sun_distortions(g["rays"], ap.shape[1], n, 7)
    assert torch.equal(du, g["dist_u"]) and torch.equal(de, g["dist_e"])
    assert torch.equal(O.reflect(g["incident"].unsqueeze(1), an), g["reflected"])
    assert torch.equal(O.scatter_rays(du, de, g["reflected"]), g["scattered"])
    be, bu, t, lam = O.ray_pixel_coordinates(ap, an, g["incident"], du, de, g["target_idx"], tg, g["res"])
    for got, name in ((be, "be"), (bu, "bu"), (t, "t"), (lam, "lambert")):
        assert torch.equal(got, g[name]), name
    flux, ic, ot, bl = O.trace_rays(ap, an, g["incident"], du, de, g["target_idx"], tg, g["res"], batch_size=3)
    assert torch.equal(flux, g["flux"])
    assert torch.equal(ic, g["intercept"]) and torch.equal(ot, g["on_target"]) and torch.equal(bl, g["blocking"])
    assert torch.equal(O.bitmaps_per_target(flux, g["target_idx"], tg.n_total), g["per_target"])
    assert g["flux"].sum() > 0 and (g["lambert"][g["target_idx"] == 1] > 0).any(), "fixture must hit both target types"


def test_batch_size_does_not_change_results(golden):
    g = golden["trace_linear"]
    tg = _targets(g["field"])
    a = O.trace_rays(g["aligned_points"], g["aligned_normals"], g["incident"], g["dist_u"], g["dist_e"], g["target_idx"],
                     tg, g["res"], batch_size=1)[0]
    assert torch.equal(a, g["flux"])


def test_nurbs_forward_and_gradient(golden):
    for g in golden["nurbs"]:
        cp = g["control_points"].clone().requires_grad_(True)
        pts, nrm = O.nurbs_points_and_normals(cp, g["degrees"][0], g["degrees"][1], g["eval_points"], g["canting"],
                                              g["facet_translations"])
        assert torch.equal(pts, g["points"]) and torch.equal(nrm, g["normals"])
        ((pts * g["weight_points"]).sum() + (nrm * g["weight_normals"]).sum()).backward()
        scale = g["grad_control_points"].abs().max()
        assert (cp.grad - g["grad_control_points"]).abs().max() <= 1e-5 * scale


def test_sampler_and_sun(golden):
    for (ns, nh, ws), lists in golden["samplers"].items():
        for rank in range(ws):
            assert O.sampler_indices(ns, nh, ws, rank) == lists[rank]
    s = golden["sun_seed7"]
    du, de = O.sun_distortions(3, 5, 2, 7)
    assert torch.equal(du, s["u"]) and torch.equal(de, s["e"])
    assert torch.equal(torch.rand(3), s["next_rand"]), "global RNG side effect of get_distortions must match"


@pytest.mark.parametrize("batch_size", [100, 4])
def test_blocking_trace_bit_exact(golden, batch_size):
    """blocking_active=True: per-batch primitive filter (the LBVH's result set) + soft mask, bit for bit."""
    g = golden["blocking"]
    tg = _targets(g["field"])
    c, s, n = O.blocking_primitives(g["aligned_points"])
    assert torch.equal(c, g["corners"][..., :3]) and torch.equal(s, g["spans"][..., :3]) and torch.equal(n, g["normals"][..., :3])
    blk = dict(corners=g["corners"], spans=g["spans"], normals=g["normals"], sample_to_blocker=torch.arange(9))
    flux, ic, ot, bl = O.trace_rays(g["aligned_points"], g["aligned_normals"], g["incident"], g["dist_u"], g["dist_e"],
                                    g["target_idx"], tg, g["res"], batch_size=batch_size, blocking=blk)
    want = g[f"batch{batch_size}"]
    assert torch.equal(flux, want["flux"]) and torch.equal(bl, want["blocking"])
    assert torch.equal(ic, want["intercept"]) and torch.equal(ot, want["on_target"])
    assert want["blocking"].min() < 0.2, "fixture must contain real shadowing"


def test_flux_center_of_mass_and_crop_bit_exact():
    """The oracle's restatement of artist/flux/bitmap.py against the real reference (tests/golden/make_flux_golden.py)."""
    import os

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flux_golden.pt"), weights_only=False)
    for key, c in g.items():
        assert torch.equal(O.flux_center_of_mass(c["flux"]), c["center_of_mass"]), key
        got = O.crop_flux_around_center(c["flux"], c["target_dimensions"], c["crop"][0], c["crop"][1])
        assert torch.equal(got, c["cropped"]), key


def test_bitmap_losses_bit_exact():
    """The oracle's restatement of PixelLoss / KLDivergenceLoss against the real reference (tests/golden/make_loss_golden.py),
    forward values and the autograd gradient w.r.t. the prediction."""
    import os

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.pt"), weights_only=False)
    for key, c in g.items():
        for name, fn in (("pixel", O.pixel_loss), ("kl", O.kl_divergence_loss)):
            pred = c["prediction"].clone().requires_grad_(True)
            loss = fn(pred, c["ground_truth"])
            assert torch.equal(loss.detach(), c[name]), (key, name)
            (loss * c["weights"]).sum().backward()
            assert torch.equal(pred.grad, c[name + "_grad"]), (key, name)
