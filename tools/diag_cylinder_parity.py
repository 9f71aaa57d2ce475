"""Where the cylindrical-target coordinates differ from the oracle, per quantity (strict trig table).
Run on a GPU box: python tools/diag_cylinder_parity.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import artist_oracle as O
from tests import cases
from tests.test_gpu_trace_parity import _run_cuda

for n, ppf, rays in ((6, (16, 16), 6), (12, (40, 40), 10)):
    case = cases.make_case(n=n, points_per_facet=ppf, rays=rays, target_pattern=(1, 0, 1))
    res = (256, 256)
    be, bu, t, lam = O.ray_pixel_coordinates(case["points"], case["normals"], case["incident"], case["dist_u"],
                                             case["dist_e"], case["target_idx"], case["targets"], res)
    (flux, ic, ot, bl), (dbe, dbu, dt, dlam) = _run_cuda(case, res, trig_mode=1, debug=True)
    cyl = case["target_idx"] == 1
    both = ((lam > 0) & (dlam.cpu() > 0))[cyl]
    print(f"case n={n} ppf={ppf} rays={rays}: cylinder rays {both.numel()}, valid in both {int(both.sum())}, "
          f"valid in one only {int(((lam > 0) != (dlam.cpu() > 0))[cyl].sum())}")
    for name, a, b in (("t", dt, t), ("bu", dbu, bu), ("be", dbe, be), ("lambert", dlam, lam)):
        a = a.cpu()[cyl][both]; b = b[cyl][both]
        print(f"  {name:8s} bit-identical {float((a == b).float().mean()):.6f}  max abs diff {float((a - b).abs().max()):.3e}"
              f"  max rel {float(((a - b).abs() / b.abs().clamp_min(1e-30)).max()):.3e}")
    ref, *_ = O.trace_rays(case["points"], case["normals"], case["incident"], case["dist_u"], case["dist_e"],
                           case["target_idx"], case["targets"], res)
    print(f"  flux max err / peak (cyl rows) {float((flux.cpu() - ref)[cyl].abs().max() / ref[cyl].max()):.3e}, peak {float(ref[cyl].max()):.3f}")
