"""GPU parity of the scoring kernel: against the oracle on random pairs and against golden values produced
by the reference's own ssim_numpy / psnr_numpy (tests/golden/scoring.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import scoring_oracle as S
from gpu_common import mod

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_scoring_golden(golden_dir):
    ops = mod("ops")
    g = np.load(os.path.join(golden_dir, "scoring.npz"))
    for k in range(int(g["n"])):
        hr, sr = g[f"c{k}.hr"], g[f"c{k}.sr"]
        wss = [int(w) for w in g[f"c{k}.ws"]]
        out = ops.score_images(torch.from_numpy(sr[None]).to(DEV), torch.from_numpy(hr[None]).to(DEV), wss).cpu().numpy()[0]
        assert np.abs(out[:len(wss)] - g[f"c{k}.ssim"]).max() < 1e-5, (k, out[:len(wss)], g[f"c{k}.ssim"])
        assert abs(out[len(wss)] - float(g[f"c{k}.mse"])) < 1e-7
        want_p = float(g[f"c{k}.psnr"])
        assert (np.isinf(want_p) and np.isinf(out[len(wss) + 1])) or abs(out[len(wss) + 1] - want_p) < 1e-4


@pytest.mark.parametrize("H,W,C", [(128, 128, 3), (128, 128, 1), (64, 96, 3)])
def test_scoring_vs_oracle_full_sweep(H, W, C):
    ops = mod("ops")
    rng = np.random.default_rng(H + C)
    B = 5
    hr = rng.integers(0, 256, size=(B, H, W, C), dtype=np.uint8)
    sr = np.clip(hr.astype(np.int32) + rng.integers(-20, 21, size=hr.shape), 0, 255).astype(np.uint8)
    sr[1] = hr[1]                              # identical pair: ssim 1, mse 0, psnr inf
    sr[2] = 0                                  # constant image
    wss = S.window_sizes_for(min(H, W))
    out = ops.score_images(torch.from_numpy(sr).to(DEV), torch.from_numpy(hr).to(DEV), wss).cpu().numpy()
    ssim, mse, psnr = S.score_images(list(sr), list(hr), wss)
    assert np.abs(out[:, :len(wss)] - ssim).max() < 1e-6
    assert np.abs(out[:, len(wss)] - mse).max() < 1e-7
    fin = np.isfinite(psnr)
    assert np.array_equal(np.isfinite(out[:, len(wss) + 1]), fin)
    assert np.abs(out[fin, len(wss) + 1] - psnr[fin]).max() < 1e-4
    assert abs(out[1, 0] - 1.0) < 1e-9 and out[1, len(wss)] == 0.0


def test_scoring_rejects_bad_window():
    ops = mod("ops")
    z = torch.zeros(1, 16, 16, 1, dtype=torch.uint8, device=DEV)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        ops.score_images(z, z, [40])
    with pytest.raises(RuntimeError, match="unsupported shape"):
        ops.score_images(z, z, [4])
