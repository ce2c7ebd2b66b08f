"""GPU end-to-end parity of the DRCT forward pass: against the oracle on the same seeded weights and inputs and
against the committed outputs of the reference itself (tests/golden/drct_*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import drct_oracle as O
from gpu_common import DrctOpt, mod

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-2          # max |dSR| / rgb_range, bf16 storage + fp32 accumulate (north star)


def _model(cfg: O.DrctCfg, sd):
    drct = mod("drct")
    opt = DrctOpt(img_size=cfg.img_size, n_colors=cfg.n_colors, embed_dim=cfg.embed_dim, layers=cfg.num_layers,
                  heads=cfg.num_heads, upscale=cfg.upscale)
    opt.window_size = cfg.window_size
    m = drct.DRCT(opt)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return m.to(DEV).eval()


def test_drct_small_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "drct_small.npz"))
    cfg = O.DrctCfg(img_size=16, n_colors=1, embed_dim=60, num_layers=4, num_heads=6, window_size=4)
    sd = O.make_state_dict(cfg, seed=3, affine_jitter=0.1)
    assert abs(O.state_dict_checksum(sd) - float(g["checksum"])) < 1e-6 * float(g["checksum"])
    m = _model(cfg, sd)
    x = torch.from_numpy(g["x"]).to(DEV)
    sr = m(x).cpu().numpy()
    err = np.abs(sr - g["sr"]).max() / 255.0
    assert err < TOL, f"max|dSR|/rgb_range = {err}"


def test_drct_l_rgb_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "drct_l_rgb.npz"))
    cfg = O.DrctCfg()
    sd = O.make_state_dict(cfg, seed=1)
    m = _model(cfg, sd)
    sr, u8 = m.run(torch.from_numpy(g["x"]).to(DEV), want_float=True, want_u8=True)
    err = np.abs(sr.cpu().numpy() - g["sr"]).max() / 255.0
    assert err < TOL, f"max|dSR|/rgb_range = {err}"
    from oracle import scoring_oracle as S
    assert np.array_equal(u8.cpu().numpy(), S.quantize_u8(sr.cpu().numpy(), 255.0))


@pytest.mark.parametrize("B", [1, 5])
def test_drct_l_vs_oracle_batch(B):
    cfg = O.DrctCfg()
    sd = O.make_state_dict(cfg, seed=2, affine_jitter=0.05)
    m = _model(cfg, sd)
    gen = torch.Generator().manual_seed(B)
    x = torch.rand(B, 3, 32, 32, generator=gen) * 255.0
    with torch.no_grad():
        want = O.drct_forward(sd, x, cfg)
    got = m(x.to(DEV)).cpu()
    err = float((got - want).abs().max()) / 255.0
    assert err < TOL, f"max|dSR|/rgb_range = {err}"
    # batch independence: image i of a batch equals the same image run alone (units are independent)
    if B > 1:
        solo = m(x[2:3].to(DEV)).cpu()
        assert torch.equal(solo[0], got[2])


def test_drct_rectangular_input_and_gray():
    cfg = O.DrctCfg(img_size=32, n_colors=1, embed_dim=60, num_layers=2, num_heads=6, window_size=8)
    sd = O.make_state_dict(cfg, seed=5, affine_jitter=0.1)
    m = _model(cfg, sd)
    x = torch.rand(2, 1, 16, 40, generator=torch.Generator().manual_seed(0)) * 255.0   # != opt.img_size: mask recomputed
    with torch.no_grad():
        want = O.drct_forward(sd, x, cfg)
    got = m(x.to(DEV)).cpu()
    assert float((got - want).abs().max()) / 255.0 < TOL


@pytest.mark.timeout(600)
def test_drct_l_64px_lr_vs_oracle():
    """BASELINE configs[3]: DRCT-L x4 at 64 px LR -> 256 px HR: 16 x 16 windows (N = 256), shift 8, 4096 tokens per image."""
    cfg = O.DrctCfg(img_size=64, window_size=16)
    sd = O.make_state_dict(cfg, seed=4, affine_jitter=0.05)
    m = _model(cfg, sd)
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(11)) * 255.0
    with torch.no_grad():
        want = O.drct_forward(sd, x, cfg)
    got = m(x.to(DEV)).cpu()
    assert got.shape == (2, 3, 256, 256)
    err = float((got - want).abs().max()) / 255.0
    assert err < TOL, f"max|dSR|/rgb_range = {err}"


def test_drct_rgb_range_1_full_model():
    """rgb_range = 1 through the whole model (both input scalings are run): inputs in [0, 1], the uint8
    truncation multiplies by 255 / rgb_range (src/evaluate.py:214)."""
    from oracle import scoring_oracle as S
    cfg = O.DrctCfg(num_layers=2)
    sd = O.make_state_dict(cfg, seed=8, affine_jitter=0.05)
    drct = mod("drct")
    opt = DrctOpt(layers=2, rgb_range=1)
    m = drct.DRCT(opt)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x = torch.rand(3, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = O.drct_forward(sd, x, cfg)
    sr, u8 = m.run(x.to(DEV), want_float=True, want_u8=True)
    err = float((sr.cpu() - want).abs().max()) / 1.0
    assert err < TOL, f"max|dSR|/rgb_range = {err}"
    assert np.array_equal(u8.cpu().numpy(), S.quantize_u8(sr.cpu().numpy(), 1.0))


def test_drct_odd_window_count_falls_back_inside_forward():
    """B * nW odd (one image of 24 x 24 px = 9 windows): the fused attention half needs window pairs, so the forward takes the
    separate qkv / attention / proj kernels for every block -- same result as the oracle."""
    cfg = O.DrctCfg(img_size=24, n_colors=3, embed_dim=180, num_layers=1, num_heads=6, window_size=8)
    sd = O.make_state_dict(cfg, seed=12, affine_jitter=0.05)
    m = _model(cfg, sd)
    ops = mod("ops")
    x = torch.rand(1, 3, 24, 24, generator=torch.Generator().manual_seed(5)) * 255.0
    with torch.no_grad():
        want = O.drct_forward(sd, x, cfg)
    ops.PROFILE = []
    try:
        got = m(x.to(DEV)).cpu()
        kinds = [p[0] for p in ops.PROFILE]
    finally:
        ops.PROFILE = None
    assert "swin_attn" not in kinds and kinds.count("window_attention") == 5
    assert float((got - want).abs().max()) / 255.0 < TOL
    x2 = torch.cat([x, x.flip(-1)])                        # two images: 18 windows, the fused kernel runs again
    ops.PROFILE = []
    try:
        got2 = m(x2.to(DEV)).cpu()
        kinds2 = [p[0] for p in ops.PROFILE]
    finally:
        ops.PROFILE = None
    assert kinds2.count("swin_attn") == 5
    assert float((got2[0] - want[0]).abs().max()) / 255.0 < TOL
