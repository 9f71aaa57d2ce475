"""How far the blocked trace is from the oracle (flux, blocking factors, primitive normals).  GPU box:
python tools/diag_blocking_parity.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import artist_oracle as O
from tests import cases
from tests.test_gpu_blocking import _scene, _oracle_blocking
from artist_b200 import HeliostatRayTracer, ops

for res, ppf, rays in (((48, 48), (10, 10), 4), ((128, 96), (10, 10), 4), ((128, 128), (24, 24), 8)):
    ft, scenario, group, mask, tidx, inc = _scene(ppf=ppf, rays=rays)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=True, bitmap_resolution=torch.tensor(res))
    flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx)
    pts, nrm = group.active_surface_points.cpu(), group.active_surface_normals.cpu()
    tg = cases.targets_from(ft)
    du, de = tracer.distortions_dataset.distortions_u.cpu(), tracer.distortions_dataset.distortions_e.cpu()
    ob = _oracle_blocking(pts, 9)
    ref, ric, rot, rbl = O.trace_rays(pts, nrm, inc.cpu(), du, de, tidx.cpu(), tg, res, blocking=ob)
    bi = tracer._blocking_inputs(tidx)
    print(f"res {res} ppf {ppf} rays {rays}: flux err/peak {float((flux.cpu() - ref).abs().max() / ref.max()):.3e}  "
          f"blocking factor err {float((bl.cpu() - rbl).abs().max()):.3e}  intercept err {float((ic.cpu() - ric).abs().max()):.3e}  "
          f"prim normals CUDA==CPU {torch.equal(bi.normals.cpu()[..., :3], ob['normals'][..., :3])} "
          f"spans {torch.equal(bi.spans.cpu()[..., :3], ob['spans'][..., :3])}")
