"""Generate ``tests/golden/flux_golden.pt``: the REAL reference's ``artist.flux.bitmap.get_center_of_mass`` and
``crop_flux_distributions_around_center`` (imported from ``/root/reference``, build container only) on small seeded
flux-like bitmaps with planar and cylindrical target areas.  Pins the oracle's restatement (tests/test_oracle_golden.py)
and is compared with the CUDA kernels directly (tests/test_gpu_flux.py).  Re-run: python tests/golden/make_flux_golden.py"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tools.ref_import import import_reference  # noqa: E402


def blobs(n: int, height: int, width: int, seed: int) -> torch.Tensor:
    """Off-centre anisotropic Gaussian spots + a little speckle: what traced flux bitmaps look like."""
    g = torch.Generator().manual_seed(seed)
    v, u = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32), indexing="ij")
    out = []
    for _ in range(n):
        cu, cv = (0.25 + 0.5 * torch.rand(2, generator=g)) * torch.tensor([width, height])
        su, sv = 1.5 + 4.0 * torch.rand(2, generator=g)
        b = (1 + 9 * torch.rand(1, generator=g)) * torch.exp(-0.5 * (((u - cu) / su) ** 2 + ((v - cv) / sv) ** 2))
        out.append(b + 0.01 * torch.rand(height, width, generator=g))
    return torch.stack(out)


def main() -> None:
    import_reference()
    from artist.field.heliostat_field import HeliostatField  # noqa: F401  (import order, see make_scenario_golden.py)
    from artist.field.solar_tower import SolarTower
    from artist.field.tower_target_areas_cylindrical import TowerTargetAreasCylindrical
    from artist.field.tower_target_areas_planar import TowerTargetAreasPlanar
    from artist.flux import bitmap

    cpu = torch.device("cpu")
    planar = TowerTargetAreasPlanar(names=["a", "b"], centers=torch.zeros(2, 4), normals=torch.zeros(2, 4),
                                    dimensions=torch.tensor([[8.0, 7.0], [5.4, 6.4]]))
    cyl = TowerTargetAreasCylindrical(names=["c"], centers=torch.zeros(1, 4), normals=torch.zeros(1, 4), axes=torch.zeros(1, 4),
                                      radii=torch.tensor([4.14]), heights=torch.tensor([5.229]),
                                      opening_angles=torch.tensor([1.0472]))
    tower = SolarTower(target_areas=[planar, cyl], device=cpu)
    out = {}
    for key, (n, h, w, seed, crop) in {"small": (5, 24, 32, 1, (6, 6)), "square": (3, 64, 64, 2, (3.0, 2.5))}.items():
        flux = blobs(n, h, w, seed)
        tidx = torch.tensor([0, 1, 2, 0, 1][:n])
        dims = torch.stack([torch.tensor([8.0, 7.0]), torch.tensor([5.4, 6.4]),
                            torch.stack([torch.tensor(4.14) * torch.tensor(1.0472), torch.tensor(5.229)])])[tidx]   # fp32 product, as the reference forms it
        out[key] = dict(flux=flux, target_idx=tidx, target_dimensions=dims, crop=crop,
                        center_of_mass=bitmap.get_center_of_mass(flux, device=cpu),
                        cropped=bitmap.crop_flux_distributions_around_center(flux, tower, tidx, crop_width=crop[0],
                                                                             crop_height=crop[1], device=cpu))
        print(key, out[key]["center_of_mass"][0].tolist(), float(out[key]["cropped"].max()))
    dst = os.path.join(ROOT, "tests", "golden", "flux_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
