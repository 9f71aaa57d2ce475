from __future__ import annotations

import torch

from ..util.env import get_device
from .kinematics_rigid_body import Kinematics


class HeliostatGroup:
    """SoA of heliostats sharing one kinematics/actuator type (``artist/field/heliostat_group.py:10-315``)."""

    def __init__(self, names: list[str], positions: torch.Tensor, surface_points: torch.Tensor,
                 surface_normals: torch.Tensor, canting: torch.Tensor, facet_translations: torch.Tensor,
                 initial_orientations: torch.Tensor, nurbs_control_points: torch.Tensor, nurbs_degrees: torch.Tensor,
                 device: torch.device | None = None) -> None:
        device = get_device(device)
        self.number_of_heliostats = len(names)
        self.number_of_facets_per_heliostat = nurbs_control_points.shape[1]
        self.names = names
        self.positions = positions
        self.surface_points = surface_points
        self.surface_normals = surface_normals
        self.canting = canting
        self.facet_translations = facet_translations
        self.initial_orientations = initial_orientations
        self.nurbs_control_points = nurbs_control_points
        self.nurbs_degrees = nurbs_degrees
        self.kinematics = Kinematics()
        self.number_of_active_heliostats = 0
        self.active_heliostats_mask = torch.zeros(self.number_of_heliostats, device=device)
        self._asp, self._asn, self._pending_alignment, self._pending_gather = surface_points, surface_normals, None, None
        self.active_canting = canting
        self.active_facet_translations = facet_translations
        self.active_nurbs_control_points = nurbs_control_points
        self._reflection_inputs = None
        self._active_rows = None

    # ``active_surface_points`` / ``active_surface_normals`` are LAZY in two ways (same values and autograd graph as the
    # eager reference whenever somebody reads them):
    #  * after ``activate_heliostats`` with a selecting / replicating mask they are a pending GATHER of the group's rows
    #    (``ops.ActivationMap``; the reference copies them with repeat_interleave, ``heliostat_group.py:256-315``);
    #  * after an alignment (``heliostat_group_rigid_body.py:217-222,265-270``) they are a pending ROTATION
    #    (rows, orientations).
    # The ray tracer consumes both pending forms directly: its kernels read surface row ``rows[sample]`` and rotate it
    # themselves (``ab200_trace_args::src_rows`` / ``::orientations``), so neither the replicated nor the aligned
    # ``[N,P,4]`` tensors ever exist in HBM on that path.
    def _materialise(self, which: int) -> None:
        from .. import ops

        if self._pending_alignment is not None:
            points, normals, orientations, amap = self._pending_alignment
            if amap is not None:
                points, normals = points.index_select(0, amap.rows.long()), normals.index_select(0, amap.rows.long())
            self._asp, self._asn = ops.align_surfaces(points, normals, orientations)
            self._pending_alignment = None
        elif self._pending_gather is not None:
            rows = self._pending_gather.rows.long()
            if which == 0:
                self._asp = self.surface_points.index_select(0, rows)
            else:
                self._asn = self.surface_normals.index_select(0, rows)
            if self._asp is not None and self._asn is not None:
                self._pending_gather = None

    @property
    def active_surface_points(self) -> torch.Tensor:
        if self._asp is None:
            self._materialise(0)
        return self._asp

    @active_surface_points.setter
    def active_surface_points(self, value: torch.Tensor) -> None:
        if self._asn is None and self._pending_alignment is not None:
            self._materialise(1)
        self._asp = value
        self._pending_alignment = None
        if self._asn is not None:
            self._pending_gather = None

    @property
    def active_surface_normals(self) -> torch.Tensor:
        if self._asn is None:
            self._materialise(1)
        return self._asn

    @active_surface_normals.setter
    def active_surface_normals(self, value: torch.Tensor) -> None:
        if self._asp is None and self._pending_alignment is not None:
            self._materialise(0)
        self._asn = value
        self._pending_alignment = None
        if self._asp is not None:
            self._pending_gather = None

    def _set_active_surface(self, points: torch.Tensor, normals: torch.Tensor) -> None:
        self._asp, self._asn, self._pending_alignment, self._pending_gather = points, normals, None, None

    def _set_pending_gather(self, amap) -> None:
        self._asp, self._asn, self._pending_alignment, self._pending_gather = None, None, None, amap

    def _record_alignment(self, orientations: torch.Tensor) -> None:
        """``rows @ O^T`` recorded, not executed (see above)."""
        if self._pending_gather is not None and self._asp is None and self._asn is None:
            pending = (self.surface_points, self.surface_normals, orientations, self._pending_gather)
        else:
            pending = (self.active_surface_points, self.active_surface_normals, orientations, None)
        self._asp, self._asn, self._pending_alignment, self._pending_gather = None, None, pending, None

    def _set_pending_alignment(self, points: torch.Tensor, normals: torch.Tensor, orientations: torch.Tensor) -> None:
        self._asp, self._asn, self._pending_alignment, self._pending_gather = None, None, (points, normals, orientations, None), None

    def _fused_alignment(self, with_map: bool = False):
        """``(points, normals, orientations)`` if the alignment has not been materialised, else None.  With ``with_map``
        a fourth entry is the pending activation map (or None) and ``points`` / ``normals`` are then the group's
        un-replicated rows; without it a pending map is resolved first (per-sample rows, one gather)."""
        if self._pending_alignment is None or self._asp is not None or self._asn is not None:
            return None
        points, normals, orientations, amap = self._pending_alignment
        if with_map:
            return points, normals, orientations, amap
        if amap is not None:
            points, normals = points.index_select(0, amap.rows.long()), normals.index_select(0, amap.rows.long())
            self._pending_alignment = (points, normals, orientations, None)
        return points, normals, orientations

    def _active_points_per_heliostat(self) -> int:
        if self._pending_alignment is not None and self._asp is None:
            return int(self._pending_alignment[0].shape[1])
        if self._pending_gather is not None and self._asp is None:
            return int(self.surface_points.shape[1])
        return int(self.active_surface_points.shape[1])

    # The reference materialises ``preferred_reflection_directions`` ([N,P,4]) inside trace_rays; the fused
    # kernel never needs it in memory, so it is evaluated only if somebody reads the attribute.
    @property
    def preferred_reflection_directions(self) -> torch.Tensor:
        if self._reflection_inputs is None:
            return torch.empty(self.number_of_heliostats, 4, device=self.positions.device)
        incident = self._reflection_inputs.unsqueeze(1)
        normals = self.active_surface_normals
        return incident - 2 * torch.sum(incident * normals, dim=-1, keepdim=True) * normals

    @preferred_reflection_directions.setter
    def preferred_reflection_directions(self, value) -> None:
        self._reflection_inputs = None
        self.__dict__["_explicit_reflection"] = value

    def align_surfaces_with_incident_ray_directions(self, aim_points, incident_ray_directions, active_heliostats_mask,
                                                    device=None) -> None:
        raise NotImplementedError("Must be overridden!")

    def align_surfaces_with_motor_positions(self, motor_positions, active_heliostats_mask, device=None) -> None:
        raise NotImplementedError("Must be overridden!")

    def activate_heliostats(self, active_heliostats_mask: torch.Tensor | None = None,
                            device: torch.device | None = None) -> None:
        """Select (and replicate: mask values > 1) heliostats (``:225-315``).  With an all-ones mask the
        ``active_*`` tensors alias the group's tensors (no copies); otherwise rows are gathered once."""
        device = get_device(device) if device is not None else self.positions.device
        if active_heliostats_mask is None:
            active_heliostats_mask = torch.ones(self.number_of_heliostats, dtype=torch.int32, device=device)
        key = (id(active_heliostats_mask), active_heliostats_mask._version)
        if getattr(self, "_mask_key", None) == key:   # same mask tensor as last time: skip the two host syncs
            n_active, identity = self._mask_info
        else:
            n_active = int(active_heliostats_mask.sum().item())
            identity = n_active == self.number_of_heliostats and bool((active_heliostats_mask == 1).all().item())
            self._mask_key, self._mask_info, self._mask_ref = key, (n_active, identity), active_heliostats_mask
        self.number_of_active_heliostats = n_active
        self.active_heliostats_mask = active_heliostats_mask
        kin = self.kinematics
        act = getattr(kin, "actuators", None)
        if identity:
            self._active_rows = None
            pick = lambda t: t
            self._set_active_surface(self.surface_points, self.surface_normals)
        else:
            from .. import ops

            if getattr(self, "_amap_key", None) != key:
                self._amap, self._amap_key = ops.ActivationMap.from_mask(active_heliostats_mask), key
            rows = self._amap.rows.long()
            self._active_rows = rows
            pick = lambda t: t.index_select(0, rows.to(t.device))
            # the [N,P,4] surfaces are NOT copied: pending gather, consumed by the ray tracer as an index map
            self._set_pending_gather(self._amap)
        self.active_canting = pick(self.canting)
        self.active_facet_translations = pick(self.facet_translations)
        self.active_nurbs_control_points = pick(self.nurbs_control_points)
        kin.number_of_active_heliostats = self.number_of_active_heliostats
        if hasattr(kin, "heliostat_positions"):
            kin.active_heliostat_positions = pick(kin.heliostat_positions)
            kin.active_initial_orientations = pick(kin.initial_orientations)
            kin.active_translation_deviation_parameters = pick(kin.translation_deviation_parameters)
            kin.active_rotation_deviation_parameters = pick(kin.rotation_deviation_parameters)
            kin.active_motor_positions = pick(kin.motor_positions)
        if act is not None:
            act.active_non_optimizable_parameters = pick(act.non_optimizable_parameters)
            if act.optimizable_parameters.numel() > 0:
                act.active_optimizable_parameters = pick(act.optimizable_parameters)
            else:
                act.active_optimizable_parameters = torch.tensor([], requires_grad=True)
