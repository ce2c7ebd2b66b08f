"""GPU parity: index maps (bit-exact), LayerNorm, window shuffles, head/tail convolutions, quantisation."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import drct_oracle as O
from oracle import scoring_oracle as S
from gpu_common import bf16_round, mod, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("H,ws", [(16, 4), (32, 8), (64, 16), (24, 4)])
def test_index_maps_bit_exact(H, ws, golden_dir):
    ops = mod("ops")
    g = np.load(os.path.join(golden_dir, "index_maps.npz"))
    for shift in (0, ws // 2):
        src, reg = ops.window_index_map(H, H, ws, shift, DEV)
        assert np.array_equal(src.cpu().numpy().reshape(-1, ws * ws), g[f"src_H{H}_ws{ws}_s{shift}"])
        assert torch.equal(src.cpu().long().view(-1, ws * ws), O.window_source_index(H, H, ws, shift))
        ids = reg.cpu().long().view(-1, ws * ws)
        if shift:
            assert torch.equal(ids, O.shift_region_ids(H, H, ws, shift))
            mask = torch.where(ids[:, None, :] != ids[:, :, None], -100.0, 0.0).numpy()
            assert np.array_equal(mask, g[f"mask_H{H}_ws{ws}"])
        else:
            assert int(ids.abs().sum()) == 0


@pytest.mark.parametrize("C", [60, 180, 212, 244, 276, 308])
def test_layernorm_rows(C):
    ops = mod("ops")
    torch.manual_seed(C)
    M, ld = 517, 320
    x = torch.zeros(M, ld, device=DEV, dtype=torch.bfloat16)
    x[:, :C] = (torch.randn(M, C, device=DEV) * 3 + 1).to(torch.bfloat16)
    x[:, C:] = 77.0                      # neighbours of the slab must not leak into the statistics
    g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    out = torch.full((M, ld), 5.0, device=DEV, dtype=torch.bfloat16)
    ops.layernorm_rows(x, out, g, b, C)
    want = F.layer_norm(x[:, :C].float(), (C,), g, b, 1e-5)
    assert (out[:, :C].float() - want).abs().max() < 0.03
    cpad = (C + 15) // 16 * 16
    if cpad > C:
        assert float(out[:, C:cpad].float().abs().max()) == 0.0
    if ld > cpad:
        assert float((out[:, cpad:].float() - 5.0).abs().max()) == 0.0


@pytest.mark.parametrize("H,ws,shift", [(32, 8, 0), (32, 8, 4), (16, 4, 2)])
def test_ln_shift_partition_and_reverse(H, ws, shift):
    ops = mod("ops")
    torch.manual_seed(1)
    B, C, ld = 3, 180, 192
    M = B * H * H
    x = torch.zeros(M, ld, device=DEV, dtype=torch.bfloat16)
    x[:, :C] = torch.randn(M, C, device=DEV).to(torch.bfloat16)
    g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    win = torch.zeros(M, ld, device=DEV, dtype=torch.bfloat16)
    ops.ln_shift_partition(x, win, g, b, B, H, H, C, ws, shift)
    src = O.window_source_index(H, H, ws, shift).reshape(-1).to(DEV)
    ln = F.layer_norm(x[:, :C].float(), (C,), g, b, 1e-5).view(B, H * H, C)
    want = ln[:, src, :].reshape(M, C)
    assert (win[:, :C].float() - want).abs().max() < 0.03
    # the reverse kernel is a pure permutation: bit-exact round trip
    back = torch.zeros(M, ld, device=DEV, dtype=torch.bfloat16)
    ops.window_reverse_unshift(win, back, B, H, H, C, ws, shift)
    back_want = torch.empty(B, H * H, C, device=DEV, dtype=torch.bfloat16)
    back_want[:, src, :] = win[:, :C].view(B, H * H, C)
    assert torch.equal(back[:, :C], back_want.view(M, C))


@pytest.mark.parametrize("nc,C", [(3, 180), (1, 60)])
def test_drct_head(nc, C):
    ops = mod("ops")
    torch.manual_seed(2)
    B, H = 2, 16
    x = torch.rand(B, nc, H, H, device=DEV) * 255
    w, bias = torch.randn(C, nc, 3, 3, device=DEV) * 0.2, torch.randn(C, device=DEV)
    mean = torch.tensor(O.RGB_MEAN if nc == 3 else (0.0,), device=DEV)
    g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    e16 = (C + 15) // 16 * 16
    x0 = torch.zeros(B * H * H, e16, device=DEV, dtype=torch.bfloat16)
    slab = torch.zeros(B * H * H, 320, device=DEV, dtype=torch.bfloat16)
    ops.drct_head(x, w, bias, mean, 1.0, g, b, C, x0, slab)
    f = F.conv2d(x - mean.view(1, nc, 1, 1), w, bias, padding=1)
    t = f.flatten(2).transpose(1, 2).reshape(-1, C)
    assert rel_err(x0[:, :C], t) < 0.01
    assert (slab[:, :C].float() - F.layer_norm(t, (C,), g, b, 1e-5)).abs().max() < 0.03


@pytest.mark.parametrize("nc,rgb_range", [(3, 255.0), (1, 255.0), (3, 1.0)])
def test_conv_last_quant(nc, rgb_range):
    ops = mod("ops")
    torch.manual_seed(3)
    B, H, Cin = 2, 24, 64
    x = (torch.randn(B * H * H, Cin, device=DEV) * 2).to(torch.bfloat16)
    w, bias = torch.randn(nc, Cin, 3, 3, device=DEV) * 0.05 * rgb_range, torch.randn(nc, device=DEV) * rgb_range * 0.3
    mean = torch.tensor(O.RGB_MEAN if nc == 3 else (0.0,), device=DEV)
    out = torch.empty(B, nc, H, H, device=DEV)
    u8 = torch.empty(B, H, H, nc, device=DEV, dtype=torch.uint8)
    ops.conv_last_quant(x, B, H, H, Cin, w, bias, nc, mean, 1.0, rgb_range, out, u8)
    xin = x.float().view(B, H, H, Cin).permute(0, 3, 1, 2)
    want = F.conv2d(xin, w, bias, padding=1) + mean.view(1, nc, 1, 1)
    assert (out - want).abs().max() < 1e-4 * float(want.abs().max())
    # the uint8 image must be the truncation of the kernel's own fp32 output (src/evaluate.py:214)
    want_u8 = S.quantize_u8(out.cpu().numpy(), rgb_range)
    assert np.array_equal(u8.cpu().numpy(), want_u8)


@pytest.mark.parametrize("B,H,W", [(2, 128, 128), (1, 40, 72), (3, 10, 18)])
def test_conv_last_quant_sizes(B, H, W):
    """Full 8 x 64 tiles (the DRCT-L HR size), ragged tiles, and a width that is not a multiple of 4 (generic kernel)."""
    ops = mod("ops")
    torch.manual_seed(H + W)
    nc, Cin, rgb_range = 3, 64, 255.0
    x = (torch.randn(B * H * W, Cin, device=DEV) * 2).to(torch.bfloat16)
    w, bias = torch.randn(nc, Cin, 3, 3, device=DEV) * 0.05 * rgb_range, torch.randn(nc, device=DEV) * rgb_range * 0.3
    mean = torch.tensor(O.RGB_MEAN, device=DEV)
    out = torch.empty(B, nc, H, W, device=DEV)
    u8 = torch.empty(B, H, W, nc, device=DEV, dtype=torch.uint8)
    ops.conv_last_quant(x, B, H, W, Cin, w, bias, nc, mean, 1.0, rgb_range, out, u8)
    xin = x.float().view(B, H, W, Cin).permute(0, 3, 1, 2)
    want = F.conv2d(xin, w, bias, padding=1) + mean.view(1, nc, 1, 1)
    assert (out - want).abs().max() < 1e-4 * float(want.abs().max())
    assert np.array_equal(u8.cpu().numpy(), S.quantize_u8(out.cpu().numpy(), rgb_range))
    u8_only = torch.empty_like(u8)
    ops.conv_last_quant(x, B, H, W, Cin, w, bias, nc, mean, 1.0, rgb_range, None, u8_only)     # the evaluator's call: no fp32 image
    assert torch.equal(u8_only, u8)


def test_quantize_u8_matches_oracle():
    ops = mod("ops")
    torch.manual_seed(4)
    x = torch.rand(3, 3, 20, 28, device=DEV) * 300 - 20
    got = ops.quantize_u8(x, 255.0).cpu().numpy()
    assert np.array_equal(got, S.quantize_u8(x.cpu().numpy(), 255.0))
    x1 = torch.rand(2, 1, 8, 8, device=DEV)
    assert np.array_equal(ops.quantize_u8(x1, 1.0).cpu().numpy(), S.quantize_u8(x1.cpu().numpy(), 1.0))
