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


def test_pack_conv3x3_compact_twin_layout():
    """K-concatenated weight image of the halo-tile conv (csrc/conv_halo.cu): K index = tap * round16(Cin) + channel, one N tile of
    round16(Cout) rows in 64-column slabs with the 128-byte swizzle; absent when the weights cannot stay resident in shared memory."""
    pack = importlib.import_module(PKG + ".pack")
    ops = importlib.import_module(PKG + ".ops")
    assert ops.halo_parts(64, 64) == 4 * 33 and ops.halo_parts(32, 32) == 4 * 9      # 4 row quadrants per 128-position tile
    torch.manual_seed(3)
    w, b = torch.randn(80, 80, 3, 3), torch.randn(80)
    pw = pack.pack_conv3x3_weight(w, b)
    c = pw.compact
    assert c is not None and c.BN == 80 and c.n_tiles == 1 and c.k_stages == 12 and c.N == 80 and c.K == 80
    assert torch.equal(c.bias[:80], b) and c.data.numel() == 12 * 80 * 128
    img = c.data.view(torch.bfloat16).view(1, c.k_stages, c.BN, 8, 8)
    wb = w.to(torch.bfloat16)
    for (n, ci, ky, kx) in [(0, 0, 0, 0), (79, 79, 2, 2), (17, 64, 1, 0), (5, 3, 0, 2), (42, 31, 2, 1)]:
        k = (ky * 3 + kx) * 80 + ci
        s, kk = divmod(k, 64)
        chunk, e = divmod(kk, 8)
        assert img[0, s, n, chunk ^ (n % 8), e] == wb[n, ci, ky, kx]
    # K tail of the last slab (720 .. 767) is zero
    flat = torch.zeros(80, 768)
    for s in range(12):
        for chunk in range(8):
            rows = torch.arange(80)
            flat[rows, s * 64 + chunk * 8:s * 64 + chunk * 8 + 8] = img[0, s, rows, chunk ^ (rows % 8)].float()
    assert flat[:, 720:].abs().sum() == 0
    assert torch.equal(flat[:, :720].view(80, 9, 80), wb.permute(0, 2, 3, 1).reshape(80, 9, 80).float())
    # channel padding to 16 per tap
    c2 = pack.pack_conv3x3_weight(torch.randn(48, 40, 3, 3), None).compact
    assert c2 is not None and c2.BN == 48 and c2.k_stages == (9 * 48 + 63) // 64
    # too wide to stay resident / too many channels: no twin, the streaming kernel is used
    assert pack.pack_conv3x3_weight(torch.randn(256, 64, 3, 3), None).compact is None
    assert pack.pack_conv3x3_weight(torch.randn(64, 180, 3, 3), None).compact is None


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


def test_pack_swin_attn_layout():
    """Fused attention half: qkv slabs per (head, K slab) with rows q_h | k_h | v_h (gamma folded, pad rows zero), the folded
    bias / column sums in the same row order, proj slabs per head with K = the head's channels."""
    pack = importlib.import_module(PKG + ".pack")
    torch.manual_seed(1)
    c, heads = 212, 4                                    # head_dim 53 -> 64, 4 K slabs
    hd, hdp, ks, cp = 53, 64, 4, 224
    qkv_w, qkv_b = torch.randn(3 * c, c), torch.randn(3 * c)
    gamma, beta = torch.rand(c) + 0.5, torch.randn(c)
    proj_w, proj_b = torch.randn(c, c), torch.randn(c)
    pa = pack.pack_swin_attn(qkv_w, qkv_b, gamma, beta, 1e-5, proj_w, proj_b, heads)
    assert (pa.C, pa.heads, pa.hd, pa.hdp) == (c, heads, hd, hdp)
    img = pa.w1.view(torch.bfloat16).view(heads, ks, 3 * hdp, 8, 8)          # [head, K slab, row, chunk position, element]
    wg = (qkv_w * gamma[None, :]).to(torch.bfloat16)
    for (h, which, d, k) in [(0, 0, 0, 0), (1, 1, 52, 200), (3, 2, 17, 64), (2, 0, 30, 211)]:
        row = which * hdp + d
        s_, kk = divmod(k, 64)
        chunk, e = divmod(kk, 8)
        assert img[h, s_, row, chunk ^ (row % 8), e] == wg[which * c + h * hd + d, k]
    assert img[:, :, hd:hdp].abs().sum() == 0 and img[:, :, hdp + hd:2 * hdp].abs().sum() == 0      # head padding rows
    t = qkv_w @ beta + qkv_b
    assert torch.allclose(pa.bias_qkv.view(heads, 3, hdp)[1, 2, 5], t[2 * c + 1 * hd + 5])
    assert pa.bias_qkv.view(heads, 3, hdp)[:, :, hd:].abs().sum() == 0
    cs = wg.float().sum(1)
    assert torch.allclose(pa.colsum_qkv.view(heads, 3, hdp)[3, 0, 7], cs[3 * hd + 7], rtol=1e-5)
    pimg = pa.w2.view(torch.bfloat16).view(heads, cp, 8, 8)
    pw = proj_w.to(torch.bfloat16)
    for (h, n, d) in [(0, 0, 0), (2, 211, 52), (3, 100, 9)]:
        chunk, e = divmod(d, 8)
        assert pimg[h, n, chunk ^ (n % 8), e] == pw[n, h * hd + d]
    assert pimg[:, c:].abs().sum() == 0 and torch.equal(pa.bias_p[:c], proj_b) and pa.bias_p[c:].abs().sum() == 0


def test_swin_mlp_plans_fit_the_sm():
    """Static tilings of the fused MLP (plain, and with the adjust conv folded into fc2) for the DRCT-L widths: TMEM columns,
    shared memory and chunk widths stay inside the limits the kernel checks."""
    pack = importlib.import_module(PKG + ".pack")
    for c, h in [(180, 360), (212, 424), (244, 488), (276, 276), (308, 308)]:
        for fuse in (False, True):
            pl = pack.swin_mlp_plan(c, h, fuse)
            n2, hc, nc = pl["n2"], pl["hc"], pl["nc"]
            assert n2 == (32 if fuse else (c + 15) // 16 * 16) and pl["fold"] == int(fuse) and pl["acc1_col"] == ((2 if fuse else 1) * n2, (2 if fuse else 1) * n2 + hc)
            assert sum(pl["widths"]) >= h and all(w % 16 == 0 and w <= hc for w in pl["widths"]) and pl["acc1_col"][1] + hc <= 512
            assert sum(pl["pieces"]) == n2 and all(16 <= r <= 256 and r % 16 == 0 for r in pl["pieces"])
            assert not fuse or (pl["pieces"] == [32] and pl["w2_slot_bytes"] == 4096)
            smem = 2 * pl["ks1"] * 16384 + pl["w1_slots"] * pl["w1_slot_bytes"] + pl["w2_slots"] * pl["w2_slot_bytes"] + \
                pack._mlp_fixed_bytes(nc * hc, n2) + ((pl["ks1"] * 4096 + 8192) if fuse else 0)
            assert smem <= 232448 and pl["w1_slots"] >= 2 and pl["w2_slots"] >= 2 and nc * hc <= 640


def test_swin_mlp_folded_adjust_pack():
    """Adjust conv folded into fc2 (W_adj W2 as the fc2 weights): the packed slabs and the bias are the products they should be."""
    pack = importlib.import_module(PKG + ".pack")
    torch.manual_seed(3)
    w2, b2, wa, ba = torch.randn(180, 360), torch.randn(180), torch.randn(32, 180, 1, 1), torch.randn(32)
    args = (torch.randn(360, 180), torch.randn(360), torch.ones(180), torch.zeros(180), 1e-5, w2, b2)
    pm, pm0 = pack.pack_swin_mlp(*args, wa, ba), pack.pack_swin_mlp(*args)
    assert pm.plan.numel() == 24 and pm.plan.tolist()[4] == 32 and pm.plan.tolist()[23] == 1
    assert pm0.plan.tolist()[23] == 0 and pm0.plan.tolist()[4] == 192 and pm0.wadj is None
    assert pm.wadj.numel() == 3 * 32 * 128 and pm.bias_adj.numel() == 32
    assert torch.allclose(pm.bias_adj, ba + wa.view(32, 180) @ b2, atol=1e-4)
    wf = (wa.view(32, 180).double() @ (0.5 * w2).double()).float().to(torch.bfloat16)
    img = pm.w2.view(torch.bfloat16).view(-1, 32, 8, 8)           # slabs of [32 rows x 64 hidden columns], (chunk, K slab) order
    nc, hc = pm.plan.tolist()[2], pm.plan.tolist()[3]
    assert nc * hc >= 360 and img.shape[0] == sum((w + 63) // 64 for w in pm.plan.tolist()[14:14 + nc])
    for (n, k) in [(0, 0), (5, 77), (31, 200)]:
        j, r = divmod(k, hc)
        slab = sum((w + 63) // 64 for w in pm.plan.tolist()[14:14 + j]) + r // 64
        chunk, e = divmod(r % 64, 8)
        assert img[slab, n, chunk ^ (n % 8), e] == wf[n, k]


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


def test_host_side_pack_functions_match_torch_restatement():
    """adsr_pack_slab_sw128 / adsr_pack_tiles_sw128 (host code in the C ABI, SURVEY.md 8b `pack_weights_*`): bit-identical to
    the gather-based torch restatement of the K-major 128-byte-swizzle layout, zero padding included."""
    import importlib
    import torch
    pack = importlib.import_module("anomaly-detection-super-resolution_b200.pack")
    torch.manual_seed(0)
    pos = torch.arange(8)
    for rows in (16, 37, 240):
        blk = torch.randn(rows, 64)
        got = pack._swizzle_slab(blk)
        t = blk.to(torch.bfloat16).view(rows, 8, 8)
        r = torch.arange(rows)
        src = (pos[None, :] ^ (r[:, None] % 8))[:, :, None].expand(rows, 8, 8)
        assert torch.equal(got, torch.gather(t, 1, src).contiguous().view(torch.uint8).reshape(-1))
    for (n, k, bn, nt) in ((200, 192, 128, 2), (32, 320, 32, 1), (540, 180, 128, 5)):
        kpad = (k + 63) // 64 * 64
        w = torch.zeros(n, kpad)
        w[:, :k] = torch.randn(n, k)
        data, ks = pack._swizzle_tiles(w, bn, nt)
        wp = torch.zeros(nt * bn, kpad)
        wp[:n] = w
        t = wp.to(torch.bfloat16).view(nt, bn, ks, 8, 8).permute(0, 2, 1, 3, 4)
        rr = torch.arange(bn)
        idx = (pos[None, :] ^ (rr[:, None] % 8))[None, None, :, :, None].expand(nt, ks, bn, 8, 8)
        assert torch.equal(data, torch.gather(t.contiguous(), 3, idx).contiguous().view(torch.uint8).reshape(-1))


def test_workspace_queries():
    import importlib
    abi = importlib.import_module("anomaly-detection-super-resolution_b200._abi")
    lib = abi.lib()
    b1 = lib.adsr_drct_workspace_bytes(1, 32, 32, 180, 32, 4, 768, 320)
    b256 = lib.adsr_drct_workspace_bytes(256, 32, 32, 180, 32, 4, 768, 320)
    assert b1 > 0 and b256 == 256 * b1                       # linear in the batch
    assert lib.adsr_drct_workspace_bytes(1, 32, 32, 180, 32, 3, 768, 320) == -1
    assert lib.adsr_score_workspace_bytes(256, 128, 128, 3, 13) == 0


def test_drn_rejects_unaligned_channel_slices_up_front():
    """The reference's x8 DRN setting (n_feats = 10) would need padded channel slices in the cat buffers: rejected at construction
    with a clear message instead of failing with ADSR_ERR_BAD_ALIGN in the middle of a forward."""
    import importlib
    import pytest as _pytest
    drn = importlib.import_module("anomaly-detection-super-resolution_b200.drn")

    class Opt:
        scale, n_blocks, n_feats, n_colors, rgb_range, negval = [2, 4, 8], 2, 10, 3, 255, 0.2

    with _pytest.raises(ValueError, match="n_feats=10"):
        drn.DRN(Opt())
    Opt.scale, Opt.n_feats = [2, 4], 20
    assert sum(p.numel() for p in drn.DRN(Opt()).parameters()) > 0
