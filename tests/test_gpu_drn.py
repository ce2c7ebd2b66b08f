"""GPU parity of the DRN path: DRN-only kernels against torch fp32, and the full DRN forward against the reference's own
outputs (tests/golden/drn_*.npz) and the oracle."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import drn_oracle as DO
from gpu_common import PKG, mod, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-2


class DrnOpt:
    def __init__(self, cfg: DO.DrnCfg):
        self.scale = [2 ** (i + 1) for i in range(cfg.phase)]
        self.n_blocks, self.n_feats, self.n_colors = cfg.n_blocks, cfg.n_feats, cfg.n_colors
        self.rgb_range, self.negval = 255, cfg.negval


def _model(cfg, sd):
    drn = mod("drn")
    m = drn.DRN(DrnOpt(cfg))
    assert set(m.state_dict().keys()) == set(sd.keys())
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return m.to(DEV).eval()


@pytest.mark.parametrize("nc,scale", [(3, 4), (1, 4), (3, 2)])
def test_bicubic_affine(nc, scale):
    ops = mod("ops")
    torch.manual_seed(nc)
    x = torch.rand(2, nc, 12, 20, device=DEV) * 255
    mat = torch.eye(nc, device=DEV) + 0.01 * torch.randn(nc, nc, device=DEV)
    bias = torch.randn(nc, device=DEV) * 10
    out = torch.empty(2, nc, 12 * scale, 20 * scale, device=DEV)
    ops.bicubic_affine(x, scale, mat.contiguous(), bias, out)
    want = F.conv2d(F.interpolate(x, scale_factor=scale, mode="bicubic", align_corners=False), mat.view(nc, nc, 1, 1), bias)
    assert (out - want).abs().max() < 2e-3


def test_conv3x3_small_two_destinations():
    ops = mod("ops")
    torch.manual_seed(0)
    x = torch.randn(2, 3, 16, 24, device=DEV)
    w, b = torch.randn(20, 3, 3, 3, device=DEV) * 0.3, torch.randn(20, device=DEV)
    o1 = torch.full((2 * 16 * 24, 32), 7.0, device=DEV, dtype=torch.bfloat16)
    o2 = torch.full((2 * 16 * 24, 48), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.conv3x3_small(x, w, b, 20, o1, o2, 20)
    want = F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1).reshape(-1, 20)
    assert rel_err(o1[:, :20], want) < 0.01 and rel_err(o2[:, 20:40], want) < 0.01
    assert float(o1[:, 20:].float().abs().max()) == 0.0
    assert float((o2[:, :20].float() - 7).abs().max()) == 0.0 and float((o2[:, 40:].float() - 7).abs().max()) == 0.0


def test_channel_mean_and_ca_scale():
    ops = mod("ops")
    torch.manual_seed(1)
    B, HW, C, Cr = 3, 1024, 80, 5
    t2 = torch.randn(B * HW, C, device=DEV).to(torch.bfloat16)
    x = torch.randn(B * HW, C, device=DEV).to(torch.bfloat16)
    mean = torch.empty(B, C, device=DEV)
    ops.channel_mean(t2, B, HW, C, mean)
    want_mean = t2.float().view(B, HW, C).mean(1)
    assert (mean - want_mean).abs().max() < 1e-4
    w1, b1 = torch.randn(Cr, C, device=DEV) * 0.3, torch.randn(Cr, device=DEV)
    w2, b2 = torch.randn(C, Cr, device=DEV) * 0.3, torch.randn(C, device=DEV)
    out = torch.empty_like(x)
    ops.rcab_ca_scale(t2, x, out, mean, w1, b1, w2, b2, B, HW, C, Cr)
    s = torch.sigmoid(F.relu(want_mean @ w1.t() + b1) @ w2.t() + b2)
    want = t2.float().view(B, HW, C) * s[:, None, :] + x.float().view(B, HW, C)
    assert rel_err(out.view(B, HW, C), want) < 0.01


def test_conv_into_channel_slice():
    ops, pack = mod("ops"), mod("pack")
    torch.manual_seed(2)
    B, H, Cin, Cout = 2, 16, 20, 40
    x = torch.zeros(B * H * H, 32, device=DEV, dtype=torch.bfloat16)
    x[:, :Cin] = torch.randn(B * H * H, Cin, device=DEV).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV) * 0.1
    pw = pack.pack_conv3x3_weight(w, None)
    out = torch.full((B * H * H, 80), 3.0, device=DEV, dtype=torch.bfloat16)
    ops.conv3x3(x, B, H, H, Cin, pw, out, ocol0=40, n_store=40)
    xin = x[:, :Cin].float().view(B, H, H, Cin).permute(0, 3, 1, 2)
    want = F.conv2d(xin, w.to(torch.bfloat16).float(), None, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert rel_err(out[:, 40:80], want) < 0.012
    assert float((out[:, :40].float() - 3).abs().max()) == 0.0
    # the slice is also a valid conv INPUT (view with a 16-byte aligned column offset)
    pw2 = pack.pack_conv3x3_weight(torch.randn(40, 40, 3, 3, device=DEV) * 0.1, None)
    o2 = torch.zeros(B * (H // 2) ** 2, 48, device=DEV, dtype=torch.bfloat16)
    ops.conv3x3(out[:, 40:], B, H, H, 40, pw2, o2, stride=2)
    assert torch.isfinite(o2.float()).all()


def test_gemm_partial_store_20_columns():
    ops, pack = mod("ops"), mod("pack")
    torch.manual_seed(3)
    a = torch.randn(300, 80, device=DEV).to(torch.bfloat16)
    w, b = torch.randn(20, 80, device=DEV) * 0.1, torch.randn(20, device=DEV)
    out = torch.full((300, 48), 5.0, device=DEV, dtype=torch.bfloat16)
    ops.tc_gemm(a, 80, pack.pack_gemm_weight(w, b), out, n_store=20)
    want = a.float() @ w.to(torch.bfloat16).float().t() + b
    assert rel_err(out[:, :20], want) < 0.012
    assert float((out[:, 20:].float() - 5).abs().max()) == 0.0


@pytest.mark.parametrize("name,cfg", [("drn_small_gray", DO.DrnCfg(n_blocks=3, n_colors=1)), ("drn_l_rgb", DO.DrnCfg())])
def test_drn_golden(golden_dir, name, cfg):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    sd = DO.make_state_dict(cfg, seed=5)
    assert abs(DO.state_dict_checksum(sd) - float(g["checksum"])) < 1e-6 * float(g["checksum"])
    m = _model(cfg, sd)
    outs = m(torch.from_numpy(g["x"]).to(DEV))
    assert len(outs) == cfg.phase + 1
    for i, o in enumerate(outs):
        err = np.abs(o.cpu().numpy() - g[f"sr{i}"]).max() / 255.0
        assert err < TOL, f"output {i}: max|dSR|/rgb_range = {err}"


def test_drn_batch_vs_oracle_and_u8():
    from oracle import scoring_oracle as S
    cfg = DO.DrnCfg(n_blocks=6)
    sd = DO.make_state_dict(cfg, seed=9)
    m = _model(cfg, sd)
    x = torch.rand(5, 3, 32, 32, generator=torch.Generator().manual_seed(1)) * 255.0
    with torch.no_grad():
        want = DO.drn_forward(sd, x, cfg)
    outs, u8 = m.run(x.to(DEV), want_float=True, want_u8=True)
    for o, wv in zip(outs, want):
        assert float((o.cpu() - wv).abs().max()) / 255.0 < TOL
    assert np.array_equal(u8.cpu().numpy(), S.quantize_u8(outs[-1].cpu().numpy(), 255.0))
