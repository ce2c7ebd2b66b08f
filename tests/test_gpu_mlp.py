"""GPU parity of the fused Swin MLP kernel (norm2 + fc1 + GELU + fc2 + residual, csrc/swin_mlp.cu) against fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

from gpu_common import mod, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mlp_case(M, C, H, seed=0):
    ops, pack = mod("ops"), mod("pack")
    torch.manual_seed(seed)
    ld = 320
    y = torch.full((M, ld), 5.0, device=DEV, dtype=torch.bfloat16)            # poison beyond C
    y[:, :C] = (torch.randn(M, C, device=DEV) * 1.5 + 0.3).to(torch.bfloat16)
    z = torch.full((M, ld), -7.0, device=DEV, dtype=torch.bfloat16)
    w1 = torch.randn(H, C, device=DEV) * 0.08
    b1 = torch.randn(H, device=DEV) * 0.2
    w2 = torch.randn(C, H, device=DEV) * 0.08
    b2 = torch.randn(C, device=DEV) * 0.2
    gamma = 1.0 + 0.2 * torch.randn(C, device=DEV)
    beta = 0.1 * torch.randn(C, device=DEV)
    pm = pack.pack_swin_mlp(w1, b1, gamma, beta, 1e-5, w2, b2)
    yf = y[:, :C].float()
    stats = torch.zeros(M, 3, 2, device=DEV)                                  # two partial slots + one unused
    half = C // 2
    stats[:, 0, 0], stats[:, 0, 1] = yf[:, :half].sum(1), (yf[:, :half] ** 2).sum(1)
    stats[:, 1, 0], stats[:, 1, 1] = yf[:, half:].sum(1), (yf[:, half:] ** 2).sum(1)
    stats[:, 2] = 1e9
    ops.swin_mlp(y, C, pm, z, stats_in=(stats, 2))
    torch.cuda.synchronize()
    z_rev = torch.full((M, ld), -7.0, device=DEV, dtype=torch.bfloat16)     # tiles walked from the last one down: same bits
    ops.swin_mlp(y, C, pm, z_rev, stats_in=(stats, 2), reverse=True)
    torch.cuda.synchronize()
    assert torch.equal(z, z_rev), "reverse tile order must not change the result"
    want = yf + F.linear(F.gelu(F.linear(F.layer_norm(yf, (C,), gamma, beta, 1e-5), w1, b1)), w2, b2)
    err = rel_err(z[:, :C], want)
    assert err < 0.012, f"M={M} C={C} H={H}: rel err {err}"
    # the TMA store clips at the tensor-map width C rounded up to a 16-byte chunk: pad columns inside that chunk receive
    # exact zeros (zero weight rows, zero bias, zero-filled residual), everything beyond stays untouched
    c8 = (C + 7) // 8 * 8
    tail = z[:, C:c8].float()
    assert bool(((tail == 0.0) | (tail == -7.0)).all()), "pad columns must be zero or untouched"
    assert float((z[:, c8:].float() + 7.0).abs().max()) == 0.0, "wrote past the 16-byte chunk holding column C-1"


@pytest.mark.parametrize("C,H", [(180, 360), (212, 424), (244, 488), (276, 276), (308, 308)])
def test_mlp_drct_shapes(C, H):
    _mlp_case(128 * 5 + 37, C, H, seed=C)


def test_mlp_single_tile():
    _mlp_case(128, 180, 360)


def test_mlp_small_dims():
    _mlp_case(300, 60, 120)        # DRCT-small widths: one hidden chunk, one K panel
    _mlp_case(300, 92, 184)


def test_mlp_many_tiles_per_cta():
    _mlp_case(128 * 148 * 3 + 5, 180, 360, seed=3)


@pytest.mark.parametrize("C,H,M", [(180, 360, 677), (212, 424, 677), (244, 488, 677), (276, 276, 677), (288, 288, 677), (60, 120, 677),
                                   (180, 360, 128), (180, 360, 128 * 148 * 3 + 5)])
def test_mlp_fused_adjust(C, H, M):
    """Adjust 1x1 conv (src/drct.py:389-393) folded into fc2: slab[:, C:C+32] = LReLU_0.2(adjust(z)) computed as
    y W_adj^T + g (W_adj W2)^T + bias, z never exists; row statistics of the 32 new columns in the given slot, the next slot zeroed."""
    ops, pack = mod("ops"), mod("pack")
    torch.manual_seed(C)
    ld = 320
    y = torch.full((M, ld), 5.0, device=DEV, dtype=torch.bfloat16)
    y[:, :C] = (torch.randn(M, C, device=DEV) * 1.5 + 0.3).to(torch.bfloat16)
    w1, b1 = torch.randn(H, C, device=DEV) * 0.08, torch.randn(H, device=DEV) * 0.2
    w2, b2 = torch.randn(C, H, device=DEV) * 0.08, torch.randn(C, device=DEV) * 0.2
    wa, ba = torch.randn(32, C, 1, 1, device=DEV) * 0.1, torch.randn(32, device=DEV) * 0.2
    gamma, beta = 1.0 + 0.2 * torch.randn(C, device=DEV), 0.1 * torch.randn(C, device=DEV)
    pm = pack.pack_swin_mlp(w1, b1, gamma, beta, 1e-5, w2, b2, wa, ba)
    yf = y[:, :C].float()
    stats = torch.zeros(M, 2, 2, device=DEV)
    stats[:, 0, 0], stats[:, 0, 1] = yf.sum(1), (yf ** 2).sum(1)
    slab = torch.full((M, ld), -7.0, device=DEV, dtype=torch.bfloat16)
    st_out = torch.full((M, 6, 2), 1e9, device=DEV)
    ops.swin_mlp_adjust(y, C, pm, slab, C, stats_in=(stats, 1), stats_out=(st_out, 2))
    torch.cuda.synchronize()
    slab_rev = torch.full((M, ld), -7.0, device=DEV, dtype=torch.bfloat16)
    ops.swin_mlp_adjust(y, C, pm, slab_rev, C, stats_in=(stats, 1), stats_out=(torch.empty_like(st_out), 2), reverse=True)
    torch.cuda.synchronize()
    assert torch.equal(slab, slab_rev), "reverse tile order must not change the result"
    z = yf + F.linear(F.gelu(F.linear(F.layer_norm(yf, (C,), gamma, beta, 1e-5), w1, b1)), w2, b2)
    want = F.leaky_relu(F.linear(z, wa.view(32, C), ba), 0.2)
    got = slab[:, C:C + 32].float()
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 0.015, f"C={C}: rel err {err}"
    outside = torch.cat([slab[:, :C], slab[:, C + 32:]], dim=1).float()
    assert float((outside + 7.0).abs().max()) == 0.0, "wrote outside the 32-column slice"
    s1, s2 = got.sum(1), (got ** 2).sum(1)
    assert float((st_out[:, 2, 0] - s1).abs().max()) < 5e-3 * float(s1.abs().max()) + 1e-3
    assert float((st_out[:, 2, 1] - s2).abs().max()) < 5e-3 * float(s2.abs().max()) + 1e-3
    assert float(st_out[:, 3].abs().max()) == 0.0
    assert float((st_out[:, :2] - 1e9).abs().max()) == 0.0 and float((st_out[:, 4:] - 1e9).abs().max()) == 0.0


@pytest.mark.parametrize("C,H,Co,M,inplace", [(308, 308, 180, 677, True), (308, 308, 180, 128 * 148 * 2 + 9, True), (244, 488, 180, 677, False),
                                              (128, 256, 60, 300, True)])
def test_mlp_folded_conv_with_residual(C, H, Co, M, inplace):
    """adjust5 + the RDG residual folded into the MLP kernel (src/drct.py:394-396): out[:, :Co] = res[:, :Co] + 0.2 (W_a z + b_a), z never
    written; in place on the slab (out is res) or into another tensor; row statistics of the Co output columns in the given slot."""
    ops, pack = mod("ops"), mod("pack")
    torch.manual_seed(C + Co)
    ld = 320
    y = torch.full((M, ld), 5.0, device=DEV, dtype=torch.bfloat16)
    y[:, :C] = (torch.randn(M, C, device=DEV) * 1.5 + 0.3).to(torch.bfloat16)
    res = torch.full((M, ld), -7.0, device=DEV, dtype=torch.bfloat16)
    res[:, :Co] = (torch.randn(M, Co, device=DEV) * 2.0).to(torch.bfloat16)
    w1, b1 = torch.randn(H, C, device=DEV) * 0.08, torch.randn(H, device=DEV) * 0.2
    w2, b2 = torch.randn(C, H, device=DEV) * 0.08, torch.randn(C, device=DEV) * 0.2
    wa, ba = torch.randn(Co, C, 1, 1, device=DEV) * 0.1, torch.randn(Co, device=DEV) * 0.2
    gamma, beta = 1.0 + 0.2 * torch.randn(C, device=DEV), 0.1 * torch.randn(C, device=DEV)
    pm = pack.pack_swin_mlp_conv_res(w1, b1, gamma, beta, 1e-5, w2, b2, wa, ba, 0.2)
    assert pm.plan.tolist()[23] == 2 and pm.plan.tolist()[4] == (Co + 15) // 16 * 16
    yf, rf = y[:, :C].float(), res[:, :Co].float()
    stats = torch.zeros(M, 2, 2, device=DEV)
    stats[:, 0, 0], stats[:, 0, 1] = yf.sum(1), (yf ** 2).sum(1)
    st_out = torch.full((M, 6, 2), 1e9, device=DEV)
    out = res if inplace else torch.full((M, ld), -3.0, device=DEV, dtype=torch.bfloat16)
    ops.swin_mlp_conv_res(y, C, pm, res, out, stats_in=(stats, 1), stats_out=(st_out, 2))
    torch.cuda.synchronize()
    z = yf + F.linear(F.gelu(F.linear(F.layer_norm(yf, (C,), gamma, beta, 1e-5), w1, b1)), w2, b2)
    want = rf + 0.2 * F.linear(z, wa.view(Co, C), ba)
    got = out[:, :Co].float()
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 0.012, f"C={C} Co={Co}: rel err {err}"
    fill = -7.0 if inplace else -3.0
    c8 = (Co + 7) // 8 * 8              # the TMA store clips at Co rounded up to a 16-byte chunk; pad columns there receive exact zeros
    tail = out[:, Co:c8].float()
    assert bool(((tail == 0.0) | (tail == fill)).all()) and float((out[:, c8:].float() - fill).abs().max()) == 0.0, "wrote outside the output columns"
    if not inplace:
        assert float((res[:, :Co].float() - rf).abs().max()) == 0.0 and float((res[:, Co:].float() + 7.0).abs().max()) == 0.0
    s1, s2 = got.sum(1), (got ** 2).sum(1)
    assert float((st_out[:, 2:6, 0].sum(1) - s1).abs().max()) < 5e-3 * float(s1.abs().max()) + 1e-2      # four partial slots
    assert float((st_out[:, 2:6, 1].sum(1) - s2).abs().max()) < 5e-3 * float(s2.abs().max()) + 1e-2
    assert float((st_out[:, :2] - 1e9).abs().max()) == 0.0
    out2 = res.clone() if inplace else torch.full((M, ld), -3.0, device=DEV, dtype=torch.bfloat16)     # reverse tile order: same bits
    if not inplace:
        ops.swin_mlp_conv_res(y, C, pm, res, out2, stats_in=(stats, 1), reverse=True)
        torch.cuda.synchronize()
        assert torch.equal(out, out2)
