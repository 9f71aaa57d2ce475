"""Device + distributed setup with the reference's contract (``artist/util/env.py:14-312``):
one process per GPU, ``env://`` rendezvous, NCCL on CUDA devices / Gloo on CPU, heliostat groups
dealt round-robin to ranks, nested sub-groups when there are more ranks than groups."""
from __future__ import annotations

import logging
from collections import defaultdict
from collections.abc import Generator
from contextlib import contextmanager
from itertools import cycle, islice
from typing import TypedDict

import torch

log = logging.getLogger(__name__)


class DdpSetup(TypedDict):
    device: torch.device
    is_distributed: bool
    is_nested: bool
    rank: int
    world_size: int
    process_subgroup: "torch.distributed.ProcessGroup | None"
    groups_to_ranks_mapping: dict[int, list[int]]
    heliostat_group_rank: int
    heliostat_group_world_size: int
    ranks_to_groups_mapping: dict[int, list[int]]


def get_device(device: torch.device | str | None = None) -> torch.device:
    """``None`` -> cuda if available else cpu (``env.py:269-312``)."""
    if device is None:
        return torch.device("cuda" if torch.cuda.is_available() else "cpu")
    return torch.device(device)


def initialize_ddp_environment(device: torch.device | None = None) -> tuple[torch.device, bool, int, int]:
    """Try ``init_process_group(env://)``; fall back to single-process (``env.py:33-93``)."""
    device = get_device(device)
    backend = "nccl" if device.type == "cuda" else "gloo"
    is_distributed, rank, world_size = False, 0, 1
    try:
        if not torch.distributed.is_initialized():
            torch.distributed.init_process_group(backend=backend, init_method="env://")
        is_distributed = True
        world_size = torch.distributed.get_world_size()
        rank = torch.distributed.get_rank()
    except Exception:
        log.info("Distributed mode disabled; running single-process.")
    if device.type == "cuda" and is_distributed:
        device = torch.device(f"cuda:{rank % torch.cuda.device_count()}")
        torch.cuda.set_device(device)
    return device, is_distributed, rank, world_size


def distribute_groups_among_ranks(world_size: int, number_of_heliostat_groups: int) -> tuple[dict[int, list[int]], bool]:
    """Round-robin groups over ranks; nested when ranks outnumber groups (``env.py:231-266``)."""
    mapping: dict[int, list[int]] = {i: [] for i in range(world_size)}
    groups = list(range(number_of_heliostat_groups))
    is_nested = world_size > number_of_heliostat_groups
    if is_nested:
        groups = list(islice(cycle(groups), world_size))
    slots = cycle(mapping.values())
    for g in groups:
        next(slots).append(g)
    return mapping, is_nested


def create_subgroups_for_nested_ddp(rank: int, groups_to_ranks_mapping: dict[int, list[int]]):
    """One process sub-group per heliostat group (``env.py:96-154``)."""
    ranks_to_groups: dict[int, list[int]] = defaultdict(list)
    for r, groups in groups_to_ranks_mapping.items():
        for g in groups:
            ranks_to_groups[g].append(r)
    group_rank, group_world, subgroup = 0, 1, None
    for _, ranks in ranks_to_groups.items():
        pg = torch.distributed.new_group(ranks=ranks)
        if rank in ranks:
            group_rank, group_world, subgroup = ranks.index(rank), len(ranks), pg
    return group_rank, group_world, subgroup, ranks_to_groups


@contextmanager
def setup_distributed_environment(number_of_heliostat_groups: int,
                                  device: torch.device | None = None) -> Generator[DdpSetup, None, None]:
    """Context manager yielding the ``DdpSetup`` dictionary (``env.py:157-228``)."""
    device, is_distributed, rank, world_size = initialize_ddp_environment(device)
    mapping, is_nested = distribute_groups_among_ranks(world_size, number_of_heliostat_groups)
    if is_nested:
        g_rank, g_world, subgroup, ranks_to_groups = create_subgroups_for_nested_ddp(rank, mapping)
    else:
        g_rank, g_world, subgroup = 0, 1, None
        ranks_to_groups = defaultdict(list)
        for r, groups in mapping.items():
            for g in groups:
                ranks_to_groups[g].append(r)
    try:
        yield DdpSetup(device=device, is_distributed=is_distributed, is_nested=is_nested, rank=rank,
                       world_size=world_size, process_subgroup=subgroup, groups_to_ranks_mapping=mapping,
                       heliostat_group_rank=g_rank, heliostat_group_world_size=g_world,
                       ranks_to_groups_mapping=ranks_to_groups)
    finally:
        if is_distributed:
            try:
                if torch.distributed.is_initialized():
                    torch.distributed.destroy_process_group()
            except Exception as exc:  # pragma: no cover
                log.error("Distributed cleanup failed: %s", exc)
