"""CPU, world_size 2 over gloo: the sharding contract of the N>1 path - every sample is owned by exactly one
rank, zero-filled partial per-target bitmaps all-reduce to the single-process result, and the autograd-aware
all-reduce gives every rank the full bitmap gradient (host-side logic of bench.py / tutorial 02)."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import torch.distributed.nn.functional as dist_fn

    from artist_b200.raytracing import RestrictedDistributedSampler
    from artist_b200.util.env import setup_distributed_environment
    from oracle import artist_oracle as O
    from tests import cases

    with setup_distributed_environment(number_of_heliostat_groups=1, device=torch.device("cpu")) as ddp:
        assert ddp["is_distributed"] and ddp["world_size"] == world and ddp["is_nested"]
        assert ddp["heliostat_group_world_size"] == world and ddp["heliostat_group_rank"] == rank
        case = cases.make_case(n=5, points_per_facet=(6, 6), rays=3)
        res = (32, 32)
        rows = RestrictedDistributedSampler(5, 5, ddp["heliostat_group_world_size"], ddp["heliostat_group_rank"]).rank_indices
        pts = case["points"].clone().requires_grad_(True)
        flux, *_ = O.trace_rays(pts, case["normals"], case["incident"], case["dist_u"], case["dist_e"], case["target_idx"],
                                case["targets"], res, sample_indices=rows)
        owned = torch.zeros(5)
        owned[rows] = 1
        assert (flux[owned == 0] == 0).all(), "rows of other ranks are zero-filled"
        local = O.bitmaps_per_target(flux, case["target_idx"], case["targets"].n_total)
        total = dist_fn.all_reduce(local, group=ddp["process_subgroup"])
        (total * total).sum().backward()
        counts = owned.clone()
        dist.all_reduce(counts)
        torch.save({"total": total.detach(), "grad": pts.grad, "rows": rows, "counts": counts},
                   os.path.join(out_dir, f"r{rank}.pt"))


def test_two_rank_sharded_flux_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import artist_oracle as O
    from tests import cases

    case = cases.make_case(n=5, points_per_facet=(6, 6), rays=3)
    pts = case["points"].clone().requires_grad_(True)
    flux, *_ = O.trace_rays(pts, case["normals"], case["incident"], case["dist_u"], case["dist_e"], case["target_idx"],
                            case["targets"], (32, 32))
    total = O.bitmaps_per_target(flux, case["target_idx"], case["targets"].n_total)
    (total * total).sum().backward()
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(world)]
    assert sorted(r[0]["rows"] + r[1]["rows"]) == [0, 1, 2, 3, 4]
    assert (r[0]["counts"] == 1).all()
    for i in range(world):
        assert torch.allclose(r[i]["total"], total, rtol=1e-6, atol=1e-6)
    # each rank holds the gradient of the rows it owns; every rank evaluates the same replicated loss, so the
    # backward of the SUM all-reduce (= all-reduce of the bitmap gradient) scales it by world_size - the reference
    # divides by world_size afterwards (aim_point_optimizer.py:698-702)
    grad = (r[0]["grad"] + r[1]["grad"]) / world
    assert torch.allclose(grad, pts.grad, rtol=1e-4, atol=1e-6 * pts.grad.abs().max().item())
