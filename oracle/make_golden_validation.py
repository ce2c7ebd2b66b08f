"""TEST INFRASTRUCTURE ONLY -- tests/golden/validation.npz from the UNMODIFIED reference.

Runs the body of the reference's validation loop `Trainer.test` (src/trainer.py:242-304) on seeded (sr, hr) pairs with the
reference's own functions: `quantize` (src/trainer.py:45-47, ROUNDING, unlike the evaluator's truncation), `calc_psnr` /
`calc_ssim` (src/trainer.py:95-99 -> src/metrics.py:70-108), one image at a time, averaged over the loader length.
The network forward is not part of this fixture (the SR tensors are seeded inputs; the forward has its own goldens).
    python oracle/make_golden_validation.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref_shim.import_reference()
    import importlib

    ref_shim.install_stubs()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    try:
        rtrainer = importlib.import_module("src.trainer")
    finally:
        sys.path.remove(ref_shim.REFERENCE_ROOT)
    out = {}
    g = torch.Generator().manual_seed(20)
    k = 0
    for (n, nc, hw, rgb_range) in ((3, 3, 128, 255.0), (3, 1, 64, 255.0), (3, 3, 64, 1.0), (3, 3, 8, 255.0)):
        hr = torch.rand(n, nc, hw, hw, generator=g) * rgb_range
        sr = hr + (torch.rand(n, nc, hw, hw, generator=g) - 0.5) * (0.15 * rgb_range)     # leaves [0, rgb_range] here and there
        sr[0] = hr[0]                                                                      # identical after rounding? not necessarily
        if rgb_range == 255.0:
            hr = hr.round()                                                                # PNG-like targets
            sr[1] = hr[1]                                                                  # identical pair -> PSNR inf for that image
        psnr, ssim = [], []
        for i in range(n):                                                                 # the loader yields batch 1
            q = rtrainer.quantize(sr[i:i + 1], rgb_range)
            psnr.append(rtrainer.calc_psnr(q, hr[i:i + 1], 4, rgb_range))
            ssim.append(rtrainer.calc_ssim(q, hr[i:i + 1], 4, rgb_range))
        out[f"c{k}.sr"], out[f"c{k}.hr"] = sr.numpy(), hr.numpy()
        out[f"c{k}.rgb_range"] = np.float64(rgb_range)
        out[f"c{k}.psnr"], out[f"c{k}.ssim"] = np.asarray(psnr, dtype=np.float64), np.asarray(ssim, dtype=np.float64)
        out[f"c{k}.eval_psnr"] = np.float64(sum(psnr) / n)                                 # eval_psnr / len(loader_test)
        out[f"c{k}.eval_ssim"] = np.float64(sum(ssim) / n)
        k += 1
    out["n"] = np.int64(k)
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "validation.npz"), **out)
    print("wrote", os.path.join(GOLD, "validation.npz"), {f"c{i}": (float(out[f'c{i}.eval_psnr']), float(out[f'c{i}.eval_ssim'])) for i in range(k)})


if __name__ == "__main__":
    main()
