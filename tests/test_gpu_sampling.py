"""-m gpu: ``Sun.get_distortions`` on CUDA (a8) - the hand-written Philox / Box-Muller kernel against
``torch.distributions.MultivariateNormal.sample`` (what the reference calls, ``artist/scene/sun.py:220-234``) on the same
device: values, generator side effects, cache."""
import time

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _upstream_sample(shape, seed, covariance=4.3681e-06, mean=0.0):
    """The reference's draw, verbatim (sun.py:118-121, 224-234)."""
    dev = torch.device(DEV)
    m = torch.tensor([mean, mean], dtype=torch.float, device=dev)
    cov = torch.tensor([[covariance, 0], [0, covariance]], dtype=torch.float, device=dev)
    dist = torch.distributions.MultivariateNormal(m, cov)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    u, e = dist.sample(shape).permute(3, 0, 1, 2)
    follow_up = (torch.rand(5, device=dev), torch.rand(5))     # the RNG streams a caller would see afterwards
    return u, e, follow_up


@pytest.mark.parametrize("shape,seed", [((1, 10, 10000), 7), ((5, 4, 2500), 0), ((3, 7, 33331), 12345),
                                        ((64, 10, 10000), 7), ((2, 1, 1), 3)])
def test_kernel_reproduces_torch_multivariate_normal_bit_for_bit(shape, seed):
    from artist_b200.scene import sun as sun_mod

    sun_mod.clear_distortion_cache()
    n, r, p = shape
    s = sun_mod.Sun(number_of_rays=r, device=torch.device(DEV))
    u, e = s.get_distortions(number_of_points=p, number_of_active_heliostats=n, random_seed=seed)
    mine_follow = (torch.rand(5, device=DEV), torch.rand(5))
    assert sun_mod._kernel_matches_torch[0], "the kernel must be the path that ran"
    ru, re_, ref_follow = _upstream_sample(shape, seed)
    assert u.shape == ru.shape == (n, r, p)
    assert torch.equal(u, ru) and torch.equal(e, re_)
    assert torch.equal(mine_follow[0], ref_follow[0]) and torch.equal(mine_follow[1], ref_follow[1])
    # the two tensors are views of ONE interleaved [N,R,P,2] buffer (the layout the trace kernels stream)
    assert e.data_ptr() == u.data_ptr() + 4 and u.stride() == (r * p * 2, p * 2, 2)


def test_non_default_sun_shape_and_mean():
    from artist_b200.scene import sun as sun_mod

    sun_mod.clear_distortion_cache()
    params = dict(distribution_type="normal", mean=1e-3, covariance=9.1e-06)
    s = sun_mod.Sun(number_of_rays=6, distribution_parameters=params, device=torch.device(DEV))
    u, e = s.get_distortions(number_of_points=4096, number_of_active_heliostats=3, random_seed=11)
    ru, re_, _ = _upstream_sample((3, 6, 4096), 11, covariance=9.1e-06, mean=1e-3)
    assert torch.equal(u, ru) and torch.equal(e, re_)


def test_cache_serves_rebuilt_tracers_and_keeps_rng_side_effects():
    from artist_b200.scene import sun as sun_mod

    sun_mod.clear_distortion_cache()
    s = sun_mod.Sun(number_of_rays=10, device=torch.device(DEV))
    u1, e1 = s.get_distortions(10000, 256, random_seed=7)
    a = torch.rand(3, device=DEV)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    u2, e2 = s.get_distortions(10000, 256, random_seed=7)
    torch.cuda.synchronize()
    warm_ms = (time.perf_counter() - t0) * 1e3
    b = torch.rand(3, device=DEV)
    assert u2.data_ptr() == u1.data_ptr() and torch.equal(a, b)
    assert warm_ms < 5.0, f"cache hit took {warm_ms:.2f} ms"
    u3, _ = s.get_distortions(10000, 256, random_seed=8)        # another seed: another sample
    assert u3.data_ptr() != u1.data_ptr() and not torch.equal(u3, u1)


def test_split_launches_above_2_pow_29_elements():
    """TensorIterator splits tensors whose byte offsets overflow int32 (numel > 2^29 floats) in halves, each half with its
    own Philox offset; the kernel's host side follows the same recursion."""
    from artist_b200.scene import sun as sun_mod

    sun_mod.clear_distortion_cache()
    shape = (27, 10, 1000003)     # 5.4e8 floats
    s = sun_mod.Sun(number_of_rays=shape[1], device=torch.device(DEV))
    u, e = s.get_distortions(shape[2], shape[0], random_seed=5)
    nxt = torch.rand(4, device=DEV)
    sun_mod.clear_distortion_cache()
    ru, re_, follow = _upstream_sample(shape, 5)
    assert torch.equal(u, ru) and torch.equal(e, re_) and torch.equal(nxt, follow[0][:4])
