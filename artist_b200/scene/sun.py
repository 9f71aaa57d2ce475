from __future__ import annotations

from typing import Any

import torch

from ..util.env import get_device
from .light_source import LightSource


class Sun(LightSource):
    """Sun with a bivariate normal sun shape (``artist/scene/sun.py:16-234``).

    ``get_distortions`` keeps the reference contract exactly - it reseeds the GLOBAL torch RNG and draws
    ``MultivariateNormal(mean, cov*I).sample((N, R, P))`` - because parity is defined on identical
    distortion samples; the two returned tensors are views of one interleaved ``[N,R,P,2]`` buffer,
    which is the layout the trace kernels stream (no repacking).
    """

    def __init__(self, number_of_rays: int,
                 distribution_parameters: dict[str, Any] = dict(distribution_type="normal", mean=0.0,
                                                                covariance=4.3681e-06),
                 device: torch.device | None = None) -> None:
        super().__init__(number_of_rays=number_of_rays)
        device = get_device(device)
        self.distribution_parameters = distribution_parameters
        if distribution_parameters["distribution_type"] != "normal":
            raise ValueError("Unknown sunlight distribution type.")
        mean = torch.tensor([distribution_parameters["mean"]] * 2, dtype=torch.float, device=device)
        cov = distribution_parameters["covariance"]
        covariance = torch.tensor([[cov, 0], [0, cov]], dtype=torch.float, device=device)
        self.distribution = torch.distributions.MultivariateNormal(mean, covariance)

    @property
    def scatter_sigma(self) -> float:
        return float(self.distribution_parameters["covariance"]) ** 0.5

    def get_distortions(self, number_of_points: int, number_of_active_heliostats: int,
                        random_seed: int = 7) -> tuple[torch.Tensor, torch.Tensor]:
        """``[N,R,P]`` distortions (u, e) - ``sun.py:199-234``.  On a CUDA device the sample is written by ONE kernel
        (``ab200_sample_distortions``) that reproduces torch's Philox ``normal_`` stream and the diagonal
        ``scale_tril`` product bit for bit (the eager path is a ``normal_`` plus one cuBLAS gemv per 65535 rays: 233 ms
        at 2048 heliostats), and repeated requests for the same ``(seed, N, R, P)`` - every optimiser of the reference
        rebuilds its tracer each epoch - are served from a cache.  Side effects are the reference's: the global CPU and
        CUDA generators are reseeded and the CUDA generator is left at the offset the eager sample would leave."""
        torch.manual_seed(random_seed)
        torch.cuda.manual_seed(random_seed)
        shape = (number_of_active_heliostats, self.number_of_rays, number_of_points)
        device = self.distribution.loc.device
        if device.type == "cuda":
            sample = _cuda_sample(self.distribution, shape, int(random_seed), device)
        else:
            sample = self.distribution.sample(shape)
        distortions_u, distortions_e = sample.permute(3, 0, 1, 2)
        return distortions_u, distortions_e


# ---- CUDA sampling: kernel + cache ------------------------------------------------------------------------------------
_sample_cache: "dict[tuple, tuple[torch.Tensor, int]]" = {}   # key -> (sample [N,R,P,2], generator offset after the draw)
_sample_cache_bytes_limit = 8 << 30
_kernel_matches_torch: dict[int, bool] = {}


def clear_distortion_cache() -> None:
    _sample_cache.clear()


def _dist_constants(distribution) -> tuple[float, float, float, float] | None:
    """(sigma_u, sigma_e, mean_u, mean_e) of a diagonal bivariate normal, or None if ``scale_tril`` is not diagonal.
    Cached on the distribution object (one device->host read per Sun)."""
    hit = distribution.__dict__.get("_ab200_constants", False)
    if hit is not False:
        return hit
    tril = distribution._unbroadcasted_scale_tril.detach().float().cpu()
    loc = distribution.loc.detach().float().cpu()
    result = None
    if tuple(tril.shape) == (2, 2) and tuple(loc.shape) == (2,) and float(tril[1, 0]) == 0.0 and float(tril[0, 1]) == 0.0:
        result = (float(tril[0, 0]), float(tril[1, 1]), float(loc[0]), float(loc[1]))
    distribution.__dict__["_ab200_constants"] = result
    return result


def _launch_sampler(out: torch.Tensor, n_pairs: int, seed: int, offset: int, consts, planar: torch.Tensor | None = None) -> int:
    import ctypes as C

    from .. import _lib, ops

    after = C.c_uint64(0)
    _lib.call("ab200_sample_distortions", out.data_ptr(), n_pairs, seed & 0xFFFFFFFFFFFFFFFF, offset, consts[0], consts[1],
              consts[2], consts[3], 0, 0, C.byref(after), None if planar is None else planar.data_ptr(), ops._stream())
    return int(after.value)


def _kernel_is_bit_exact(distribution, consts, device) -> bool:
    """Once per device and process: does the kernel reproduce this torch build's ``MultivariateNormal.sample``?  (It
    depends on torch's launch geometry and cuRAND's arithmetic; if either ever changes, sampling falls back to the
    eager distribution - slow, but the identical-samples contract is what parity is defined on.)"""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _kernel_matches_torch:
        gen = torch.cuda.default_generators[idx]
        state = gen.get_state()
        shape = (3, 4, 25013)      # odd sizes, several counters per thread
        gen.manual_seed(1234)
        want = distribution.sample(shape)
        want_offset = gen.get_offset()
        got = torch.empty(*shape, 2, device=device)
        got_offset = _launch_sampler(got, got.numel() // 2, 1234, 0, consts)
        _kernel_matches_torch[idx] = bool(torch.equal(want, got)) and want_offset == got_offset
        gen.set_state(state)
        if not _kernel_matches_torch[idx]:
            import warnings

            warnings.warn("artist_b200: ab200_sample_distortions does not reproduce this torch build's CUDA normal_ "
                          "stream; Sun.get_distortions uses the eager distribution (slow) to keep identical samples")
    return _kernel_matches_torch[idx]


def _cuda_sample(distribution, shape, seed: int, device) -> torch.Tensor:
    consts = _dist_constants(distribution)
    if consts is None or not _kernel_is_bit_exact(distribution, consts, device):
        return distribution.sample(shape)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[idx]
    gen.manual_seed(seed)   # (the self-check above may have run after the caller's reseed)
    key = (idx, seed, tuple(shape), consts)
    hit = _sample_cache.get(key)
    if hit is None:
        sample = torch.empty(*shape, 2, device=device)
        # the same samples de-interleaved ([2,N,R,P]), written by the same kernel: the layout the v3 trace kernels stream
        # (ab200_trace_args::distortions_planar); registered with ops so that the tracer finds it behind the sample
        # (only when those opt-in kernels are enabled, ops.use_planar: it doubles the sample's footprint)
        from .. import ops

        planar = torch.empty(2, *shape, device=device) if ops.use_planar else None
        n_pairs = sample.numel() // 2
        offset_after = _launch_sampler(sample, n_pairs, seed, 0, consts, planar) if n_pairs else 0
        if planar is not None:
            ops.register_planar_distortions(sample, planar)
        per = 8 if planar is not None else 4
        nbytes = sample.numel() * per
        while _sample_cache and sum(t.numel() * per for t, _ in _sample_cache.values()) + nbytes > _sample_cache_bytes_limit:
            _sample_cache.pop(next(iter(_sample_cache)))
        if nbytes <= _sample_cache_bytes_limit:
            _sample_cache[key] = (sample, offset_after)
    else:
        sample, offset_after = hit
    gen.set_offset(offset_after)
    return sample
