"""GPU: how often does device trig (sincosf / polynomial) move a ray into another pixel than torch-CPU trig?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from artist_b200 import ops
from oracle import artist_oracle as O
from tests import cases

dev = torch.device("cuda:0")
case = cases.make_case(n=16, points_per_facet=(50, 50), rays=10)
res = (256, 256)
be, bu, _, lam = O.ray_pixel_coordinates(case["points"], case["normals"], case["incident"], case["dist_u"], case["dist_e"],
                                         case["target_idx"], case["targets"], res)
f = lambda x: x.to(dev).float().contiguous()
tg = case["targets"]
tgt = ops.TargetTensors(f(tg.planar_centers), f(tg.planar_normals), f(tg.planar_dimensions), f(tg.cyl_centers), f(tg.cyl_normals),
                        f(tg.cyl_axes), f(tg.cyl_radii), f(tg.cyl_heights), f(tg.cyl_opening_angles))
dist = ops.pack_distortions(case["dist_u"].to(dev), case["dist_e"].to(dev))
x = torch.cat([case["dist_u"].reshape(-1), case["dist_e"].reshape(-1)])
for mode, name in ((0, "sincosf"), (2, "polynomial")):
    opt = ops.TraceOptions(res_e=res[0], res_u=res[1], trig_mode=mode, scatter_sigma=2.09e-3)
    _, (dbe, dbu, _, dlam) = ops.trace_debug(f(case["points"]), f(case["normals"]), f(case["incident"]), dist,
                                             case["target_idx"].to(dev), tgt, opt)
    flips = ((dbe.cpu().long() != be.long()) | (dbu.cpu().long() != bu.long())).sum().item()
    neq = ((dbe.cpu() != be) | (dbu.cpu() != bu)).sum().item()
    s, c = ops.debug_trig(x.to(dev), mode)
    ds = (s.cpu() != torch.sin(x)).float().mean().item()
    dc = (c.cpu() != torch.cos(x)).float().mean().item()
    print(f"{name:11s}: rays {be.numel()}  valid {int((lam > 0).sum())}  pixel-index flips {flips}  coordinate bit-mismatches {neq} "
          f"({neq / be.numel():.2e})  sin!=cpu {ds:.2e}  cos!=cpu {dc:.2e}")
