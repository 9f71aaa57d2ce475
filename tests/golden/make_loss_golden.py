"""Generate ``tests/golden/loss_golden.pt``: the REAL reference's ``artist.optim.loss.PixelLoss`` and ``KLDivergenceLoss``
(imported from ``/root/reference``, build container only) on small seeded flux-like bitmaps, forward values and the
autograd gradient w.r.t. the prediction.  Pins the oracle's restatement (tests/test_oracle_golden.py) and is compared with
the CUDA kernels directly (tests/test_gpu_flux.py).  Re-run: python tests/golden/make_loss_golden.py"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.golden.make_flux_golden import blobs  # noqa: E402
from tools.ref_import import import_reference  # noqa: E402


def main() -> None:
    import_reference()
    from artist.field.heliostat_field import HeliostatField  # noqa: F401  (import order, see make_scenario_golden.py)
    from artist.optim.loss import KLDivergenceLoss, PixelLoss

    out = {}
    for key, (n, h, w, seed) in {"small": (5, 24, 32, 11), "square": (3, 64, 64, 12)}.items():
        gt = blobs(n, h, w, seed)
        pred0 = blobs(n, h, w, seed + 100) * 1.7
        pred0[0, :3] = 0.0          # exact zeros in the prediction: log(0 + eps), |x| at 0
        weights = torch.linspace(0.5, 2.0, n)
        entry = dict(prediction=pred0, ground_truth=gt, weights=weights)
        for name, fn in (("pixel", PixelLoss()), ("kl", KLDivergenceLoss())):
            pred = pred0.clone().requires_grad_(True)
            loss = fn(prediction=pred, ground_truth=gt, reduction_dimensions=(1, 2))
            (loss * weights).sum().backward()
            entry[name] = loss.detach()
            entry[name + "_grad"] = pred.grad.clone()
            print(key, name, loss.detach().tolist())
        out[key] = entry
    dst = os.path.join(ROOT, "tests", "golden", "loss_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
