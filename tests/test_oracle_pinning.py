"""CPU tests: the oracle (oracle/*.py) against the committed golden vectors that were produced by
running the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import drct_oracle as O
from oracle import scoring_oracle as S


def test_index_maps_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "index_maps.npz"))
    for (H, ws) in [(16, 4), (32, 8), (64, 16), (24, 4)]:
        for shift in (0, ws // 2):
            got = O.window_source_index(H, H, ws, shift).numpy()
            assert np.array_equal(got, g[f"src_H{H}_ws{ws}_s{shift}"]), (H, ws, shift)
        assert np.array_equal(O.attention_mask(H, H, ws, ws // 2).numpy(), g[f"mask_H{H}_ws{ws}"])
        assert np.array_equal(O.relative_position_index(ws).numpy(), g[f"rpi_ws{ws}"])


def test_masked_windows_are_the_wrapped_ones():
    # non-zero mask only for windows {3,7,11,12..15} with a 4x4 window grid
    m = O.attention_mask(32, 32, 8, 4)
    nz = [w for w in range(16) if (m[w] != 0).any()]
    assert nz == [3, 7, 11, 12, 13, 14, 15]


def test_drct_small_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "drct_small.npz"))
    cfg = O.DrctCfg(img_size=16, n_colors=1, embed_dim=60, num_layers=4, num_heads=6, window_size=4)
    sd = O.make_state_dict(cfg, seed=3, affine_jitter=0.1)
    assert abs(O.state_dict_checksum(sd) - float(g["checksum"])) < 1e-6 * float(g["checksum"])
    taps = {}
    with torch.no_grad():
        y = O.drct_forward(sd, torch.from_numpy(g["x"]), cfg, taps)
    assert np.abs(y.numpy() - g["sr"]).max() < 2e-4
    assert np.abs(taps["embed"].numpy() - g["tap.embed"]).max() < 1e-4
    assert np.abs(taps["layers.0.out"].numpy() - g["tap.l0.out"]).max() < 1e-4


def test_drct_l_rgb_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "drct_l_rgb.npz"))
    cfg = O.DrctCfg()
    sd = O.make_state_dict(cfg, seed=1)
    assert abs(O.state_dict_checksum(sd) - float(g["checksum"])) < 1e-6 * float(g["checksum"])
    with torch.no_grad():
        y = O.drct_forward(sd, torch.from_numpy(g["x"]), cfg)
    assert np.abs(y.numpy() - g["sr"]).max() < 5e-4


def test_flops_match_survey():
    # 60.436 GFLOP (RGB 32->128), 287.798 GFLOP (64->256)
    assert abs(O.flops_per_image(O.DrctCfg()) / 1e9 - 60.436) < 0.01
    c4 = O.DrctCfg(img_size=64, window_size=16)
    assert abs(O.flops_per_image(c4) / 1e9 - 287.798) < 0.01


def test_scoring_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "scoring.npz"))
    for k in range(int(g["n"])):
        hr, sr = g[f"c{k}.hr"], g[f"c{k}.sr"]
        gh, gs = S.to_gray01(hr), S.to_gray01(sr)
        for ws, want in zip(g[f"c{k}.ws"], g[f"c{k}.ssim"]):
            assert abs(S.ssim_box(gh, gs, int(ws)) - want) < 2e-6, (k, ws)
        assert abs(S.mse01(sr, hr) - float(g[f"c{k}.mse"])) < 1e-9
        want_p = float(g[f"c{k}.psnr"])
        got_p = S.psnr01(hr, sr)
        assert (np.isinf(want_p) and np.isinf(got_p)) or abs(got_p - want_p) < 1e-6


def test_ssim_loops_equals_box():
    rng = np.random.default_rng(0)
    a = rng.random((12, 14)).astype(np.float32)
    b = np.clip(a + 0.05 * rng.normal(size=a.shape), 0, 1).astype(np.float32)
    for ws in (3, 5, 9):
        assert abs(S.ssim_loops(a, b, ws) - S.ssim_box(a, b, ws)) < 2e-6


def test_auc_matches_sklearn_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "auc.npz"))
    for i in range(int(g["n"])):
        assert abs(S.roc_auc(g[f"y{i}"], g[f"s{i}"]) - float(g[f"auc{i}"])) < 1e-12


def test_quantize_truncates():
    x = np.array([[[[-3.0, 0.4, 0.999, 127.999, 254.7, 255.0, 300.0]]]], dtype=np.float32)
    assert S.quantize_u8(x, 255.0).reshape(-1).tolist() == [0, 0, 0, 127, 254, 255, 255]


def test_window_sizes():
    assert S.window_sizes_for(128) == list(range(3, 124, 10))
    assert len(S.window_sizes_for(256)) == 26


def test_validation_oracle_matches_reference_goldens(golden_dir):
    """quantize (rounding) + psnr_torch + ssim_torch, the body of Trainer.test: oracle restatement vs the reference's own outputs."""
    import os
    import numpy as np
    from oracle import scoring_oracle as S
    g = np.load(os.path.join(golden_dir, "validation.npz"))
    for k in range(int(g["n"])):
        ps, ss, ep, es = S.validate_pairs(g[f"c{k}.sr"], g[f"c{k}.hr"], float(g[f"c{k}.rgb_range"]))
        want_p, want_s = g[f"c{k}.psnr"], g[f"c{k}.ssim"]
        fin = np.isfinite(want_p)
        assert np.array_equal(np.isinf(ps), np.isinf(want_p))
        assert np.abs(ps[fin] - want_p[fin]).max() < 1e-5 and np.abs(ss - want_s).max() < 1e-6
