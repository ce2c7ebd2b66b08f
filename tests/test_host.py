"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, packing layout,
state_dict drop-in parity with the reference layout (via the oracle's key/shape table)."""
import importlib
import os
import re

import pytest
import torch

PKG = "anomaly-detection-super-resolution_b200"


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module(PKG)


def test_library_loads_and_exports_all_symbols(pkg):
    abi = importlib.import_module(PKG + "._abi")
    handle = abi.lib()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "adsr_b200.h")).read()
    declared = set(re.findall(r"\b(adsr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(abi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.adsr_abi_version() == abi.ABI_VERSION
    assert handle.adsr_status_string(1).decode() == "unsupported shape"


def test_choose_bn():
    pack = importlib.import_module(PKG + ".pack")
    for n in (3, 32, 64, 180, 212, 256, 276, 308, 360, 424, 488, 576, 768, 864, 960):
        bn, nt = pack.choose_bn(n)
        assert bn % 32 == 0 and 32 <= bn <= 256 and bn * nt >= n and bn * (nt - 1) < n


def test_pack_gemm_layout_roundtrip():
    pack = importlib.import_module(PKG + ".pack")
    torch.manual_seed(0)
    w = torch.randn(40, 100)
    pw = pack.pack_gemm_weight(w, torch.arange(40.0))
    assert pw.BN == 64 and pw.n_tiles == 1 and pw.k_stages == 2
    img = pw.data.view(torch.bfloat16).view(pw.n_tiles, pw.k_stages, pw.BN, 8, 8)
    wb = w.to(torch.bfloat16)
    for (r, k) in [(0, 0), (5, 17), (39, 99), (13, 64), (7, 63)]:
        s, kk = divmod(k, 64)
        chunk, e = divmod(kk, 8)
        assert img[0, s, r, chunk ^ (r % 8), e] == wb[r, k]
    assert img[0, :, 40:].abs().sum() == 0 and pw.bias[39] == 39 and pw.bias[40:].abs().sum() == 0


def test_pack_qkv_head_padding():
    pack = importlib.import_module(PKG + ".pack")
    c, heads = 212, 4            # head_dim 53 -> padded to 64
    w = torch.randn(3 * c, c)
    b = torch.randn(3 * c)
    pw = pack.pack_qkv_weight(w, b, heads)
    assert pw.N == 3 * heads * 64 and pw.K == c
    # bias of head 1, dim 52 of the k block sits at padded row (1*heads + 1)*64 + 52
    assert pw.bias[(heads + 1) * 64 + 52] == b[c + 53 + 52]
    assert pw.bias[(heads + 1) * 64 + 53] == 0


def test_state_dict_matches_reference_layout(pkg):
    from oracle import drct_oracle as O
    drct = importlib.import_module(PKG + ".drct")

    class Opt:
        img_size, n_colors, embed_dim = 32, 3, 180
        depths = (6,) * 12
        num_heads = (6,) * 12
        window_size, mlp_ratio, upscale, img_range = 8, 2, 4, 1.0
        upsampler, resi_connection, rgb_range = "pixelshuffle", "1conv", 255
        compress_ratio, squeeze_factor, conv_scale, overlap_ratio = 3, 30, 0.01, 0.5

    m = drct.DRCT(Opt())
    sd_ref = O.make_state_dict(O.DrctCfg(), seed=1)     # key/shape table pinned to the reference by test_oracle_pinning
    sd = m.state_dict()
    assert set(sd) == set(sd_ref)
    assert len(sd) == 1000
    for k in sd:
        assert tuple(sd[k].shape) == tuple(sd_ref[k].shape), k
    res = m.load_state_dict(sd_ref, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.layers[0].swin2.attn_mask, sd_ref["layers.0.swin2.attn_mask"])
    assert torch.equal(m.layers[3].swin1.attn.relative_position_index, sd_ref["layers.3.swin1.attn.relative_position_index"])


def test_forward_refuses_cpu(pkg):
    drct = importlib.import_module(PKG + ".drct")

    class Opt:
        img_size, n_colors, embed_dim = 16, 1, 60
        depths = (6,) * 2
        num_heads = (6,) * 2
        window_size, mlp_ratio, upscale, img_range = 4, 2, 4, 1.0

    m = drct.DRCT(Opt())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 16, 16))
