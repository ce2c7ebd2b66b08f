"""GPU parity of the `src.metrics` drop-in signatures (psnr/ssim_numpy, psnr/ssim_torch) against values produced by the
reference's own functions (tests/golden/metrics_api.npz, oracle/make_golden.py)."""
import importlib
import os

import numpy as np
import pytest
import torch

from gpu_common import PKG

pytestmark = pytest.mark.gpu


def test_metrics_api_golden(golden_dir):
    m = importlib.import_module(PKG + ".metrics")
    g = np.load(os.path.join(golden_dir, "metrics_api.npz"))
    for k in range(int(g["n"])):
        hr, sr, rr = torch.from_numpy(g[f"t{k}.hr"]), torch.from_numpy(g[f"t{k}.sr"]), float(g[f"t{k}.rgb_range"])
        assert abs(m.psnr_torch(sr, hr, rr) - float(g[f"t{k}.psnr"])) < 1e-3
        assert abs(m.ssim_torch(sr, hr, rr) - float(g[f"t{k}.ssim"])) < 1e-5
        assert abs(m.ssim_torch(sr, hr, rr, win_size=7) - float(g[f"t{k}.ssim7"])) < 1e-5
        a = (hr[0].permute(1, 2, 0).numpy() / rr).astype(np.float32)
        b = (sr[0].permute(1, 2, 0).numpy() / rr).astype(np.float32)
        assert abs(m.ssim_numpy(a, b, 5) - float(g[f"t{k}.np_ssim5"])) < 1e-5
        assert abs(m.psnr_numpy(a, b) - float(g[f"t{k}.np_psnr"])) < 1e-3
        assert abs(m.ssim_numpy(a * 255, b * 255, 5, data_range=255.0) - float(g[f"t{k}.np_ssim5_dr255"])) < 1e-5


def test_src_alias_package_is_the_same_code():
    import src.metrics as sm
    m = importlib.import_module(PKG + ".metrics")
    assert sm.ssim_numpy is m.ssim_numpy and sm.psnr_torch is m.psnr_torch
