"""CPU: inline known-answer vectors of the reference's own unit tests, restated against the oracle.
Sources: tests/raytracing/test_geometry.py:13-83 (reflect), tests/raytracing/test_sampling.py:6-51,
tests/nurbs/test_surfaces.py:202-300 (NURBS forward), tests/geometry/test_transforms.py (rotate_distortions)."""
import math

import pytest
import torch

from oracle import artist_oracle as O


def test_reflect_kat():
    inc = torch.tensor([[1.0, 1.0, 1.0, 0.0], [1.0, 1.0, 1.0, 0.0], [1.0, 1.0, 1.0, 0.0], [2.0, 1.0, 3.0, 0.0]])
    nrm = torch.tensor([[0.0, 0.0, 1.0, 0.0], [0.0, 1.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0], [0.3, 0.6, 0.7, 0.0]])
    want = torch.tensor([[1.0, 1.0, -1.0, 0.0], [1.0, -1.0, 1.0, 0.0], [-1.0, 1.0, 1.0, 0.0], [0.02, -2.96, -1.62, 0.0]])
    torch.testing.assert_close(O.reflect(inc, nrm), want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("ns,nh,ws,want", [
    (12, 4, 1, [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]]),
    (12, 4, 2, [[0, 1, 2, 6, 7, 8], [3, 4, 5, 9, 10, 11]]),
    (12, 4, 3, [[0, 1, 2, 9, 10, 11], [3, 4, 5], [6, 7, 8]]),
    (12, 4, 4, [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11]]),
    (4, 1, 3, [[0, 1, 2, 3], [], []]),
    (4, 2, 3, [[0, 1], [2, 3], []]),
])
def test_sampler_kat(ns, nh, ws, want):
    from artist_b200.raytracing import RestrictedDistributedSampler

    for rank in range(ws):
        assert O.sampler_indices(ns, nh, ws, rank) == want[rank]
        assert list(RestrictedDistributedSampler(ns, nh, ws, rank)) == want[rank]


def test_nurbs_forward_kat():
    canting = torch.tensor([[[[8.0249e-01, -0.0, -4.7736e-03, 0.0], [1.7949e-05, 6.3749e-01, 3.0172e-03, 0.0]]]])
    tr = torch.tensor([[[1.0, 0.0, 0.0, 0.0]]])
    ev = torch.cartesian_prod(torch.linspace(1e-5, 1 - 1e-5, 2), torch.linspace(1e-5, 1 - 1e-5, 2))[None, None]
    cp = O.planar_control_points(4, 4, canting[0])[None]
    pts, nrm = O.nurbs_points_and_normals(cp, 2, 2, ev, canting, tr)
    want_p = torch.tensor([[[[1.975133419037e-01, -6.374730467796e-01, 1.756353536621e-03, 1.0],
                             [1.975492835045e-01, 6.374730467796e-01, 7.790592499077e-03, 1.0],
                             [1.802450656891e00, -6.374730467796e-01, -7.790592499077e-03, 1.0],
                             [1.802486538887e00, 6.374729871750e-01, -1.756352838129e-03, 1.0]]]])
    want_n = torch.tensor([0.005948313046, -0.004732967820, 0.999971091747, 0.0]).expand(1, 1, 4, 4)
    torch.testing.assert_close(pts, want_p)
    torch.testing.assert_close(nrm, want_n)


def test_rotate_distortions_structure():
    e, u = torch.tensor([[[0.3]]]), torch.tensor([[[-0.2]]])
    m = O.rotate_distortions(e=e, u=u)[0, 0, 0]
    ce, se, cu, su = math.cos(0.3), math.sin(0.3), math.cos(-0.2), math.sin(-0.2)
    want = torch.tensor([[cu, -su, 0, 0], [ce * su, ce * cu, -se, 0], [se * su, se * cu, ce, 0], [0, 0, 0, 1.0]])
    torch.testing.assert_close(m, want)
    with pytest.raises(ValueError):
        O.rotate_distortions(e=torch.zeros(1, 2, 3), u=torch.zeros(1, 2, 4))


def test_line_plane_axis_aligned_kat():
    """One ray straight at the centre of an 8x8 target facing +N: lands in the middle pixel, full cosine."""
    tg = O.Targets(planar_centers=torch.tensor([[0.0, 0.0, 50.0, 1.0]]), planar_normals=torch.tensor([[0.0, 1.0, 0.0, 0.0]]),
                   planar_dimensions=torch.tensor([[8.0, 8.0]]))
    dirs = torch.tensor([[[[0.0, -1.0, 0.0, 0.0]]]])
    origins = torch.tensor([[[1.0, 30.0, 51.0, 1.0]]])
    be, bu, t, lam = O.line_plane_intersections(dirs, torch.ones(1, 1, 1), origins, tg.planar_centers, tg.planar_normals,
                                                tg.planar_dimensions, torch.zeros(1, dtype=torch.int32), torch.tensor([256, 256]))
    torch.testing.assert_close(t, torch.tensor([[[30.0]]]))
    torch.testing.assert_close(lam, torch.tensor([[[1.0]]]))
    torch.testing.assert_close(be, torch.tensor([[[255.0 - (1.0 + 4.0) / 8.0 * 255.0]]]))
    torch.testing.assert_close(bu, torch.tensor([[[(1.0 + 4.0) / 8.0 * 255.0]]]))
    # a ray travelling away from the plane is invalid: everything zero except the flipped e coordinate
    be, bu, t, lam = O.line_plane_intersections(-dirs, torch.ones(1, 1, 1), origins, tg.planar_centers, tg.planar_normals,
                                                tg.planar_dimensions, torch.zeros(1, dtype=torch.int32), torch.tensor([256, 256]))
    assert be.item() == 255.0 and bu.item() == 0.0 and t.item() == 0.0 and lam.item() == 0.0


def test_splat_drops_last_row_and_column():
    """Rays exactly on the last pixel row/column are dropped (heliostat_ray_tracer.py:723-728)."""
    res = torch.tensor([8, 6])
    be = torch.tensor([[[7.0, 3.25, 0.0]]])
    bu = torch.tensor([[[2.0, 5.0, 1.5]]])
    out = O.bilinear_splatting(be, bu, torch.ones(1, 1, 3), res)
    assert out.shape == (1, 6, 8)
    torch.testing.assert_close(out.sum(), torch.tensor(1.0))          # only the third ray lands
    assert out[0, 6 - 1 - 1, 0] == 0.5 and out[0, 6 - 1 - 2, 0] == 0.5  # rows flipped


# the reference's known-answer tests of the bitmap losses (tests/optim/test_loss_functions.py:266-342 pixel, :345-445 KL)
PIXEL_KATS = [
    ([[[1.0, 2.0], [3.0, 4.0]]], [[[1.0, 2.0], [3.0, 4.0]]], [0.0]),
    ([[[2.0, 3.0], [9.0, 12.0]]], [[[1.0, 2.0], [8.0, 6.0]]], [2.2941176470588234]),
]
_A = [[0.5, 0.75, 0.41], [0.11, 2.55, 3.09]]
_B = [[5.4, 5.71, 2.46], [2.86, 0.44, 0.11]]
KL_KATS = [
    ([[[0.5, 0.5]]], [[[0.5, 0.5]]], [0.0]),
    ([_B, _A], [_A, _B], [2.311237096786, 1.351369142532]),
    ([_A, _B], [_B, _A], [1.351369142532, 2.311237096786]),
]


def test_bitmap_loss_kats():
    for pred, gt, want in PIXEL_KATS:
        torch.testing.assert_close(O.pixel_loss(torch.tensor(pred), torch.tensor(gt)), torch.tensor(want), atol=1e-6, rtol=1e-6)
    for pred, gt, want in KL_KATS:
        torch.testing.assert_close(O.kl_divergence_loss(torch.tensor(pred), torch.tensor(gt)), torch.tensor(want), atol=1e-6,
                                   rtol=1e-6)


def test_cpu_cross_product_rounding_the_kernels_reproduce():
    """Parity note behind `cross_comp` (csrc/common.cuh): torch's CPU cross kernel rounds a1*b2 - a2*b1 as
    fma(a1, b2, -RN(a2*b1)), and normalize() takes the norm as the FMA chain fma(z,z, fma(y,y, x*x)) with a correctly
    rounded sqrt.  The NURBS normals (surfaces.py:615-661), the canting frame (transforms.py:321-337) and the
    cylinder frame (geometry.py:299) of the CUDA kernels are built on exactly this sequence; if a torch build ever
    rounds differently, the bit-exact GPU tests of those rows fail and this test says why."""
    import numpy as np

    g = torch.Generator().manual_seed(5)
    a4, b4 = torch.randn(4, 3, 2500, 4, generator=g), torch.randn(4, 3, 2500, 4, generator=g)
    a, b = a4[..., :3], b4[..., :3]                      # strided views, as the oracle's s10[..., :3]
    f64 = lambda t: t.numpy().astype(np.float64)
    f32 = lambda x: x.astype(np.float32)

    def comp(i, j):   # fma(a_i, b_j, -RN(a_j * b_i)): product exact in double, one rounding to float
        return f32(f64(a[..., i]) * f64(b[..., j]) - f64(torch.from_numpy(f32(f64(a[..., j]) * f64(b[..., i])))))

    want = np.stack([comp(1, 2), comp(2, 0), comp(0, 1)], axis=-1)
    got = torch.linalg.cross(a, b)
    assert np.array_equal(got.numpy(), want)
    x, y, z = (want[..., k] for k in range(3))
    chain = f32(f64(torch.from_numpy(f32(f64(torch.from_numpy(f32(x.astype(np.float64) ** 2)))
                                         + y.astype(np.float64) ** 2))) + z.astype(np.float64) ** 2)
    norm = np.sqrt(chain.astype(np.float64)).astype(np.float32)
    unit = want / np.maximum(norm, np.float32(1e-12))[..., None]
    assert np.array_equal(torch.nn.functional.normalize(got, dim=3).numpy(), unit)


def _f32(x):
    import numpy as np
    return np.asarray(x, dtype=np.float32)


def _dot3_left_to_right(a, b):
    """sum over the last axis (size 3) of separately rounded products, added left to right - torch's CPU reduction."""
    p = _f32(a * b)
    return _f32(_f32(p[..., 0] + p[..., 1]) + p[..., 2])


def test_cpu_rounding_sequence_of_the_blocking_geometry():
    """Parity note behind `block_tuv_strict` (csrc/blocking_device.cuh): the reference's ray/rectangle intersection
    (blocking.py:318-346) on the CPU is exactly 'every product and sum its own rounding, component sums left to right,
    IEEE divisions'.  The numpy emulation below is the kernel's instruction sequence; it must equal torch bit for bit."""
    import numpy as np

    g = torch.Generator().manual_seed(2)
    B, R, P, K = 2, 3, 400, 5
    origins = torch.randn(B, P, 4, generator=g) * 5
    dirs = torch.nn.functional.normalize(torch.randn(B, R, P, 4, generator=g), dim=-1)
    corners = torch.randn(K, 4, 4, generator=g) * 5
    su_t, sv_t = corners[:, 1, :3] - corners[:, 0, :3], corners[:, 3, :3] - corners[:, 0, :3]
    nn_t = torch.nn.functional.normalize(torch.linalg.cross(su_t, sv_t), dim=-1)
    o = origins[:, None, :, None, :3]
    d = dirs[:, :, :, None, :3]
    c0, su, sv, nn = corners[None, None, None, :, 0, :3], su_t[None, None, None], sv_t[None, None, None], nn_t[None, None, None]
    den = torch.sum(d * nn, dim=-1)
    t = torch.sum((c0 - o) * nn, dim=-1) / den
    off = (o + t[..., None] * d) - c0
    uu, vv, uv = torch.sum(su * su, dim=-1), torch.sum(sv * sv, dim=-1), torch.sum(su * sv, dim=-1)
    pu, pv = torch.sum(off * su, dim=-1), torch.sum(off * sv, dim=-1)
    det = uu * vv - uv * uv
    u = (pu * vv - pv * uv) / det
    v = (pv * uu - pu * uv) / det

    n = lambda x: x.numpy()
    DEN = _dot3_left_to_right(n(d), n(nn))
    T = _f32(_dot3_left_to_right(_f32(n(c0) - n(o)), n(nn)) / DEN)
    OFF = _f32(_f32(n(o) + _f32(T[..., None] * n(d))) - n(c0))
    UU, VV, UV = (_dot3_left_to_right(n(a), n(b)) for a, b in ((su, su), (sv, sv), (su, sv)))
    PU, PV = _dot3_left_to_right(OFF, n(su)), _dot3_left_to_right(OFF, n(sv))
    DET = _f32(_f32(UU * VV) - _f32(UV * UV))
    U = _f32(_f32(_f32(PU * VV) - _f32(PV * UV)) / DET)
    V = _f32(_f32(_f32(PV * UU) - _f32(PU * UV)) / DET)
    for name, got, want in (("den", DEN, den), ("t", T, t), ("off", OFF, off), ("det", DET, det), ("u", U, u), ("v", V, v)):
        assert np.array_equal(got, n(want)), name


def test_cpu_rounding_sequence_of_the_cylinder_quadratic():
    """Parity note behind `hit_cylinder` (csrc/trace_device.cuh): frame rows from the contracted cross product, the two
    GEMMs as FMA chains over k, the quadratic with one rounding per product / sum (geometry.py:299-345).  Up to the
    discriminant the emulation equals torch bit for bit; the root goes through torch's CPU sqrt, which (MKL VML) is not
    the IEEE root for a fraction of a per cent of its arguments - the only reason the GPU's hit distance is not
    bit-identical for every ray."""
    import numpy as np

    g = torch.Generator().manual_seed(9)
    N, R, P = 3, 4, 300
    nn = torch.nn.functional.normalize(torch.tensor([[0.0, 0.9063, -0.4226], [0.3, 0.8, -0.2], [-0.2, 0.9, 0.1]]), dim=-1)
    ax = torch.nn.functional.normalize(torch.linalg.cross(nn, torch.tensor([[1.0, 0.0, 0.0]]).expand(3, -1)), dim=-1)
    cc = torch.tensor([[0.0, -3.0, 55.0]]).expand(3, -1) + torch.randn(3, 3, generator=g)
    o = torch.randn(N, P, 3, generator=g) * 3 + torch.tensor([10.0, 80.0, 2.0])
    d = torch.nn.functional.normalize(cc[:, None, None, :] - o[:, None, :, :] + 0.05 * torch.randn(N, R, P, 3, generator=g), dim=-1)
    rad = torch.tensor([4.14, 3.0, 5.5])
    uu = torch.cross(nn, ax, dim=-1)
    rot = torch.stack([uu, nn, ax], dim=1)
    ol = ((o - cc[:, None, :]) @ rot.transpose(1, 2))[:, None, :, :]
    dl = d @ rot.transpose(1, 2)[:, None, :, :]
    ox, oy, dx, dy = ol[..., 0], ol[..., 1], dl[..., 0], dl[..., 1]
    a = dx**2 + dy**2
    b = 2 * (ox * dx + oy * dy)
    c = (ox**2 + oy**2 - rad.view(-1, 1, 1) ** 2).repeat(1, R, 1)
    disc = b**2 - 4 * a * c

    n = lambda x: x.numpy()
    f64 = lambda x: np.asarray(x, dtype=np.float64)
    fma = lambda x, y, z: _f32(f64(x) * f64(y) + f64(z))
    chain = lambda vec, row: fma(vec[..., 2], row[..., 2], fma(vec[..., 1], row[..., 1], _f32(vec[..., 0] * row[..., 0])))
    q = _f32(n(o) - n(cc)[:, None, :])
    rows = n(rot)
    OL = np.stack([chain(q, rows[:, k, :][:, None, :]) for k in range(3)], axis=-1)[:, None]
    DL = np.stack([chain(n(d), rows[:, k, :][:, None, None, :]) for k in range(3)], axis=-1)
    assert np.array_equal(OL, n(ol)) and np.array_equal(DL, n(dl))
    OX, OY, DX, DY = OL[..., 0], OL[..., 1], DL[..., 0], DL[..., 1]
    A = _f32(_f32(DX * DX) + _f32(DY * DY))
    Bq = _f32(np.float32(2) * _f32(_f32(OX * DX) + _f32(OY * DY)))
    C = _f32(_f32(_f32(OX * OX) + _f32(OY * OY)) - _f32(n(rad) * n(rad))[:, None, None])
    DISC = _f32(_f32(Bq * Bq) - _f32(_f32(np.float32(4) * A) * C))
    assert np.array_equal(A, n(a)) and np.array_equal(Bq, n(b)) and np.array_equal(DISC, n(disc))
    assert (n(disc) > 0).mean() > 0.5, "the case must hit the cylinder"


def test_cpu_rounding_sequence_of_the_small_batched_matrix_products():
    """Parity note behind `m4_mul` / `motor_from_normal` (csrc/kinematics.cu): torch's CPU bmm of the kinematic chain's
    tiny operands ([N,4,4] @ [N,4,4], @ a broadcast [1,4,4], [N,4,4]^T @ [N,4,1]) adds separately rounded products left
    to right - unlike the large GEMMs (`points @ O^T`), which accumulate as an FMA chain."""
    import numpy as np

    g = torch.Generator().manual_seed(4)
    a, b, one = torch.randn(300, 4, 4, generator=g), torch.randn(300, 4, 4, generator=g), torch.randn(1, 4, 4, generator=g)

    def unfused(x, y):
        x, y = np.broadcast_arrays(x, y)
        out = np.zeros(x.shape, dtype=np.float32)
        for i in range(4):
            for j in range(4):
                acc = _f32(x[..., i, 0] * y[..., 0, j])
                for k in (1, 2, 3):
                    acc = _f32(acc + _f32(x[..., i, k] * y[..., k, j]))
                out[..., i, j] = acc
        return out

    assert np.array_equal(unfused(a.numpy(), b.numpy()), (a @ b).numpy())
    assert np.array_equal(unfused(a.numpy(), one.numpy()), (a @ one).numpy())
    assert np.array_equal(unfused(one.numpy(), a.numpy()), (one @ a).numpy())
    v = torch.randn(300, 4, generator=g)
    at = a.transpose(-1, -2).numpy()
    acc = _f32(at[:, :, 0] * v.numpy()[:, None, 0])
    for k in (1, 2, 3):
        acc = _f32(acc + _f32(at[:, :, k] * v.numpy()[:, None, k]))
    assert np.array_equal(acc, (a.transpose(-1, -2) @ v[:, :, None])[:, :, 0].numpy())
    south = torch.tensor([0.0, -1.0, 0.0, 0.0])
    assert np.array_equal((a @ south).numpy(), -a.numpy()[:, :, 1])


def test_cpu_cosine_rounding_bias_the_production_polynomial_reproduces():
    """Parity note behind `sincos_tiny` (csrc/common.cuh): for sun-shape angles torch's CPU cos is the correctly rounded
    float for only ~91 % of the arguments; it rounds 1 - x^2/2 up once the fraction of the last ulp exceeds 0.40 + 6|x|.
    The kernel's c = RN(1 - (x^2/2 - (0.1 - 6|x|) 2^-24)) - emulated here in float32 - is torch's value for > 99.9 %."""
    import numpy as np

    g = torch.Generator().manual_seed(3)
    x = torch.randn(1 << 18, generator=g) * (4.3681e-06 ** 0.5)
    want = torch.cos(x).numpy()
    xn = x.numpy()
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    fma = lambda a, b, c: _f32(f64(a) * f64(b) + f64(c))
    z = _f32(xn * xn)
    a, b = np.float32(5.9604645e-09), np.float32(3.5762787e-07)
    model = fma(np.float32(-1.0), fma(b, np.abs(xn), fma(np.float32(0.5), z, -a)), np.float32(1.0))
    correctly_rounded = _f32(np.cos(f64(xn)))
    assert (model != want).mean() < 1e-3, f"{(model != want).mean():.2e}"
    assert (correctly_rounded != want).mean() > 0.05      # why the correctly rounded cosine was not good enough
    sin_model = fma(_f32(xn * z), np.float32(-1.6666667163e-1), xn)
    assert (sin_model != torch.sin(x).numpy()).mean() < 1e-4
