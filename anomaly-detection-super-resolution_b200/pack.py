"""Load-time weight packing: turns reference-layout fp32 parameters into the bf16 shared-memory images
that tc_gemm.cu streams with one bulk copy per ring stage.

A packed weight is n_tiles tiles; tile t, K-stage s is a contiguous [BN rows x 64 bf16] block in the
canonical K-major 128-byte-swizzle layout of tcgen05 (row r at byte r*128, its 16-byte chunk c stored
at chunk position c ^ (r % 8)).  Rows >= N and K columns beyond the real K are exact zeros, so padded
outputs are exactly zero and padded inputs never contribute.

Packed tensors are derived buffers: they are NOT part of the state_dict and are
rebuilt whenever the parameters change.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional

import torch


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def to_device(obj, device):
    """Packed images are built on the HOST (the pack_* functions run wherever their inputs live; the models hand them CPU
    copies of the parameters) and reach the GPU with one plain copy per buffer: no device kernel runs at load time.
    Moves every tensor field of a Packed* dataclass (recursively) except the host-side `plan`."""
    import dataclasses

    if obj is None or not dataclasses.is_dataclass(obj):
        return obj
    for f in dataclasses.fields(obj):
        v = getattr(obj, f.name)
        if isinstance(v, torch.Tensor):
            if f.name != "plan":
                setattr(obj, f.name, v.to(device))
        elif dataclasses.is_dataclass(v):
            to_device(v, device)
    return obj


def choose_bn(n: int) -> tuple[int, int]:
    """(BN, n_tiles): BN a multiple of 32 (epilogue chunks of 32 columns never straddle N tiles), <= 256, and the
    (n_tiles, BN) pair with the least padded columns, each extra tile being charged like 40 padded columns (per-tile
    epilogue overhead, narrower MMAs)."""
    best = None
    t0 = (n + 255) // 256
    for n_tiles in range(t0, t0 + 3):
        bn = round_up((n + n_tiles - 1) // n_tiles, 32)
        if bn > 256:
            continue
        cand = (bn * n_tiles + 40 * n_tiles, n_tiles, bn)
        if best is None or cand < best:
            best = cand
    return best[2], best[1]


@dataclass
class PackedWeight:
    data: torch.Tensor        # uint8 [n_tiles * k_stages * BN * 128]
    bias: torch.Tensor        # fp32 [n_tiles * BN]
    colsum: Optional[torch.Tensor] = None   # fp32 [n_tiles * BN]: LayerNorm-folded weights only (sum_k gamma_k W_nk)
    ln_eps: float = 1e-5
    N: int = 0                # logical output columns
    K: int = 0                # logical K (GEMM) or Cin (conv)
    BN: int = 0
    n_tiles: int = 0
    k_stages: int = 0
    compact: Optional["PackedWeight"] = None   # 3x3 convs whose weights fit in shared memory: K-concatenated twin (conv_halo.cu)


def _swizzle_tiles(w2d: torch.Tensor, bn: int, n_tiles: int) -> tuple[torch.Tensor, int]:
    """w2d: fp32 [N, Kpad] with Kpad % 64 == 0 -> packed uint8 image."""
    n, kpad = w2d.shape
    ks = kpad // 64
    if w2d.device.type == "cpu":                       # host parameters: the library's host-side packer (adsr_pack_tiles_sw128)
        from . import _abi
        src = w2d.detach().float().contiguous()
        out = torch.empty(n_tiles * ks * bn * 128, dtype=torch.uint8)
        _abi.check(_abi.lib().adsr_pack_tiles_sw128(src.data_ptr(), src.stride(0), n, kpad, bn, n_tiles, out.data_ptr()),
                   "adsr_pack_tiles_sw128")
        return out, ks
    wp = torch.zeros(n_tiles * bn, kpad, dtype=torch.float32, device=w2d.device)
    wp[:n] = w2d
    t = wp.to(torch.bfloat16).view(n_tiles, bn, ks, 8, 8).permute(0, 2, 1, 3, 4)   # [tile, ks, row, chunk, elem]
    rows = torch.arange(bn, device=w2d.device)
    pos = torch.arange(8, device=w2d.device)
    src_chunk = pos[None, :] ^ (rows[:, None] % 8)                                 # chunk stored at position p
    idx = src_chunk[None, None, :, :, None].expand(n_tiles, ks, bn, 8, 8)
    out = torch.gather(t.contiguous(), 3, idx).contiguous()
    return out.view(torch.uint8).reshape(-1), ks


def _pad_bias(bias: Optional[torch.Tensor], n: int, total: int, device) -> torch.Tensor:
    b = torch.zeros(total, dtype=torch.float32, device=device)
    if bias is not None:
        b[:n] = bias.float()
    return b


def pack_gemm_weight(weight: torch.Tensor, bias: Optional[torch.Tensor], rows_kernel: bool = False) -> PackedWeight:
    """weight: [N, K] (nn.Linear / 1x1 conv layout).  rows_kernel=True packs N tiles of 128 rows, the layout the row-tile
    kernel (csrc/tc_gemm_rows.cu, K <= 320, 16-byte aligned output columns) streams."""
    w = weight.detach().float().reshape(weight.shape[0], -1)
    n, k = w.shape
    bn, n_tiles = (128, (n + 127) // 128) if (rows_kernel and k <= 320) else choose_bn(n)
    kpad = round_up(k, 64)
    w2 = torch.zeros(n, kpad, dtype=torch.float32, device=w.device)
    w2[:, :k] = w
    data, ks = _swizzle_tiles(w2, bn, n_tiles)
    return PackedWeight(data, _pad_bias(bias, n, bn * n_tiles, w.device), N=n, K=k, BN=bn, n_tiles=n_tiles, k_stages=ks)


def pack_ln_gemm_weight(weight: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                        eps: float = 1e-5, rows_kernel: bool = False) -> PackedWeight:
    """Linear applied to LayerNorm(x) with the normalisation folded into the GEMM:
         LN(x) W^T + b = rstd * (x (gamma*W)^T - mean * s) + t,   s_n = sum_k gamma_k W_nk,   t_n = sum_k beta_k W_nk + b_n.
    The kernel multiplies RAW rows with the packed gamma*W and applies mean / rstd / s / t in its epilogue; s is taken from
    the bf16-rounded packed weights so that the subtraction cancels exactly what the tensor core accumulated."""
    w = weight.detach().float().reshape(weight.shape[0], -1)
    wg = w * gamma.detach().float()[None, :]
    t = w @ beta.detach().float() + (bias.detach().float() if bias is not None else 0.0)
    pw = pack_gemm_weight(wg, t, rows_kernel)
    s = wg.to(torch.bfloat16).float().sum(dim=1)
    pw.colsum = _pad_bias(s, w.shape[0], pw.BN * pw.n_tiles, w.device)
    pw.ln_eps = float(eps)
    return pw


def pack_conv3x3_weight(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> PackedWeight:
    """weight: [Cout, Cin, 3, 3] -> K ordered [tap = ky*3+kx][cin padded to a multiple of 64]."""
    w = weight.detach().float()
    cout, cin = w.shape[0], w.shape[1]
    spt = (cin + 63) // 64
    bn, n_tiles = choose_bn(cout)
    w2 = torch.zeros(cout, 9, spt * 64, dtype=torch.float32, device=w.device)
    w2[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    data, ks = _swizzle_tiles(w2.reshape(cout, 9 * spt * 64), bn, n_tiles)
    pw = PackedWeight(data, _pad_bias(bias, cout, bn * n_tiles, w.device), N=cout, K=cin, BN=bn, n_tiles=n_tiles, k_stages=ks)
    # halo-tile kernel (csrc/conv_halo.cu): one N tile of round16(cout) rows, K index = tap * round16(cin) + channel, resident in
    # shared memory next to a two-panel halo tile (<= ~125 KB of weights)
    bnc, kt = round_up(cout, 16), round_up(cin, 16)
    slabs = (9 * kt + 63) // 64
    if cout <= 128 and cin <= 128 and slabs * bnc * 128 <= 126 * 1024:
        wc = torch.zeros(cout, slabs * 64, dtype=torch.float32, device=w.device)
        wc[:, :9 * kt].view(cout, 9, kt)[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, 9, cin)
        cdata, cks = _swizzle_tiles(wc, bnc, 1)
        pw.compact = PackedWeight(cdata, _pad_bias(bias, cout, bnc, w.device), N=cout, K=cin, BN=bnc, n_tiles=1, k_stages=cks)
    return pw


def head_pad(hd: int) -> int:
    return round_up(hd, 16)


def pack_qkv_weight(weight: torch.Tensor, bias: Optional[torch.Tensor], heads: int, gamma: Optional[torch.Tensor] = None,
                    beta: Optional[torch.Tensor] = None, eps: float = 1e-5) -> PackedWeight:
    """qkv Linear [3C, C] -> rows regrouped as q|k|v blocks of `heads` heads, each padded to head_pad(hd)
    rows (zero weights and bias), the layout adsr_window_attention reads."""
    w = weight.detach().float()
    c = w.shape[1]
    hd = c // heads
    hdp = head_pad(hd)
    w3 = torch.zeros(3, heads, hdp, c, dtype=torch.float32, device=w.device)
    w3[:, :, :hd] = w.view(3, heads, hd, c)
    b3 = torch.zeros(3, heads, hdp, dtype=torch.float32, device=w.device)
    if bias is not None:
        b3[:, :, :hd] = bias.detach().float().view(3, heads, hd)
    if gamma is not None:                       # norm1 folded into the qkv GEMM
        return pack_ln_gemm_weight(w3.view(3 * heads * hdp, c), b3.view(-1), gamma, beta, eps, rows_kernel=True)
    return pack_gemm_weight(w3.view(3 * heads * hdp, c), b3.view(-1), rows_kernel=True)


def pack_proj_weight(weight: torch.Tensor, bias: Optional[torch.Tensor], heads: int) -> PackedWeight:
    """proj Linear [C, C]: input columns regrouped to the head-padded attention output layout."""
    w = weight.detach().float()
    c = w.shape[1]
    hd = c // heads
    hdp = head_pad(hd)
    w2 = torch.zeros(w.shape[0], heads, hdp, dtype=torch.float32, device=w.device)
    w2[:, :, :hd] = w.view(w.shape[0], heads, hd)
    return pack_gemm_weight(w2.view(w.shape[0], heads * hdp), bias, rows_kernel=True)


# ------------------------------------------------------------------------------------------------ fused Swin MLP
_SMEM_LIMIT = 232448            # 227 KB of shared memory per CTA
def _mlp_fixed_bytes(hidden_padded: int, n2: int) -> int:
    """bias1 / colsum1 / bias2 caches + row statistics of two tiles + barriers (swin_mlp.cu)."""
    return (2 * hidden_padded + n2) * 4 + 2048 + 512


@dataclass
class PackedMlp:
    w1: torch.Tensor          # uint8: fc1 slabs [hcw[j] rows x 64 bf16], 128-byte swizzle, (chunk, K slab) order
    w2: torch.Tensor          # uint8: fc2 slabs [piece rows x 64 bf16], (chunk, K slab, piece) order
    bias1: torch.Tensor       # fp32 [nc * hc]
    colsum1: torch.Tensor     # fp32 [nc * hc]
    bias2: torch.Tensor       # fp32 [n2]
    plan: torch.Tensor        # int32 (host): see include/adsr_b200.h, adsr_swin_mlp_bf16
    ln_eps: float
    C: int
    H: int
    wadj: Optional[torch.Tensor] = None      # fused adjust conv: uint8 slabs [32 rows x 64 bf16] per K slab, 128-byte swizzle
    bias_adj: Optional[torch.Tensor] = None  # fp32 [32]
    conv_out: int = 0                        # pack_swin_mlp_conv_res: output channels of the folded conv (plan[23] = 2)


_ADJ_N = 32                     # output channels of the fusable adjust convs (gc of the RDG)


def swin_mlp_plan(c: int, h: int, fuse_adj: bool = False, conv_out: int = 0) -> dict:
    """Static tiling of one fused MLP (swin_mlp.cu): hidden chunks of <= 128 columns (two fp32 chunk accumulators plus the
    fc2 accumulator must fit the 512 TMEM columns), the fc2 output issued in two pieces of >= 128 rows when it is wider
    than 255, and the shared memory left after the two y-tile buffers split between the fc1 and fc2 weight rings."""
    # fuse_adj: the adjust conv is folded INTO fc2 (W_adj W2, 32 rows), so the "fc2" accumulator is the 32 adjust columns;
    # conv_out > 0: a wide 1x1 conv with residual folded in the same way (adsr_swin_mlp_conv_res_bf16): its output channels
    n2 = _ADJ_N if fuse_adj else round_up(conv_out if conv_out else c, 16)
    if conv_out and (fuse_adj or (n2 + 63) // 64 > (c + 63) // 64):
        raise ValueError(f"folded conv does not fit: C={c} -> {conv_out}")
    k1steps = (c + 15) // 16
    ks1 = (c + 63) // 64
    # TMEM: fc2 accumulator (n2) + two fc1 chunk accumulators (2 hc); narrower chunks when two fc1 ring slots would not fit
    a2 = 2 * n2 if fuse_adj else n2              # folded adjust: the 32-column accumulator is double-buffered
    hc_max = min(128, ((512 - a2) // 2) // 16 * 16)
    if conv_out:        # chunks narrow enough for THREE fc1 accumulators next to the wide one: smaller ring slots, a deeper fc2 ring
        hc_max = min(hc_max, ((512 - a2) // 3) // 16 * 16)
    while True:
        nc = (h + hc_max - 1) // hc_max
        hc = round_up((h + nc - 1) // nc, 16)
        widths = [hc] * (nc - 1) + [round_up(h - hc * (nc - 1), 16)]
        s1 = round_up(hc * 128, 1024)
        avail = _SMEM_LIMIT - 2 * ks1 * 16384 - _mlp_fixed_bytes(nc * hc, n2) - ((ks1 * _ADJ_N * 128 + 8192) if fuse_adj else 0)
        # the fc2 N dimension goes out in one piece, or in two when it is wider than 255 -- or when one-piece ring slots would
        # not leave room for two slots per ring
        if n2 >= 256 or (avail < 2 * s1 + 2 * round_up(n2 * 128, 1024) and n2 >= 64):
            p0 = round_up(n2 // 2, 16)
            pieces = [p0, n2 - p0]
        else:
            pieces = [n2]
        s2 = round_up(max(pieces) * 128, 1024)
        n1, n2s = 2, 2
        if avail >= n1 * s1 + n2s * s2 or hc_max <= 48:
            break
        hc_max -= 16
    if avail < n1 * s1 + n2s * s2 or nc > 8 or n2 > 320 or nc * hc > 640:
        raise ValueError(f"fused MLP does not fit: C={c} H={h}")
    # bytes per tile through each ring decide how the slots are shared out (at least 2 each)
    n2s_max = 4 if fuse_adj else 8               # the 4 KB slabs of a folded adjust: more slots only take ring space from fc1
    while True:
        grew = False
        # (a folded wide conv streams ks1 * pieces extra slabs through the fc2 ring at every tile start: that ring grows first)
        for which in ((2, 1) if conv_out and n2s < 4 else (1, 2) if n1 * s1 <= n2s * s2 else (2, 1)):
            if which == 1 and n1 < 8 and avail >= (n1 + 1) * s1 + n2s * s2:
                n1 += 1
                grew = True
                break
            if which == 2 and n2s < n2s_max and avail >= n1 * s1 + (n2s + 1) * s2:
                n2s += 1
                grew = True
                break
        if not grew:
            break
    return dict(ks1=ks1, k1steps=k1steps, nc=nc, hc=hc, n2=n2, widths=widths, pieces=pieces, w1_slots=n1, w1_slot_bytes=s1,
                w2_slots=n2s, w2_slot_bytes=s2, acc1_col=(a2, a2 + hc), adj_tcol=0, fold=2 if conv_out else int(fuse_adj))


def _swizzle_slab(block: torch.Tensor) -> torch.Tensor:
    """[rows, 64] fp32 -> uint8 image, row r at r*128 B with its 16-byte chunk c at position c ^ (r % 8)."""
    rows = block.shape[0]
    if block.device.type == "cpu":                     # host parameters: adsr_pack_slab_sw128
        from . import _abi
        src = block.detach().float().contiguous()
        out = torch.empty(rows * 128, dtype=torch.uint8)
        _abi.check(_abi.lib().adsr_pack_slab_sw128(src.data_ptr(), src.stride(0), rows, 64, out.data_ptr()), "adsr_pack_slab_sw128")
        return out
    t = block.to(torch.bfloat16).view(rows, 8, 8)
    r = torch.arange(rows, device=block.device)
    pos = torch.arange(8, device=block.device)
    src = (pos[None, :] ^ (r[:, None] % 8))[:, :, None].expand(rows, 8, 8)
    return torch.gather(t, 1, src).contiguous().view(torch.uint8).reshape(-1)


# W_adj (y + fc2(g)) = y W_adj^T + g (W_adj W2)^T: fc2 and the adjust conv have nothing non-linear between them, and z itself is
# never needed when the adjust conv is fused -- so its 32 rows replace fc2's C rows (6-10x fewer fc2 MMAs, no residual epilogue).
def pack_swin_mlp(fc1_w, fc1_b, gamma, beta, eps, fc2_w, fc2_b, adjust_w=None, adjust_b=None) -> PackedMlp:
    """norm2 + fc1 + GELU + fc2 of one Swin block (src/drct.py:438-441, 173-190) for adsr_swin_mlp_bf16: gamma folded into
    fc1 (pack_ln_gemm_weight's algebra), the 0.5 of GELU folded into fc2 (exact in bf16).  With adjust_w [32, C(,1,1)] the
    RDG's adjust 1x1 conv is packed along for adsr_swin_mlp_adjust_bf16 (raises ValueError if the tiling does not fit)."""
    w1 = fc1_w.detach().float()
    w2 = fc2_w.detach().float() * 0.5
    h, c = w1.shape
    dev = w1.device
    fuse_adj = adjust_w is not None
    if fuse_adj and adjust_w.shape[0] != _ADJ_N:
        raise ValueError("only 32-channel adjust convs can be fused")
    fold = fuse_adj
    wa32 = adjust_w.detach().float().reshape(_ADJ_N, -1) if fuse_adj else None
    pl = swin_mlp_plan(c, h, fuse_adj)
    hc, nc, n2, ks1 = pl["hc"], pl["nc"], pl["n2"], pl["ks1"]
    w1g = torch.zeros(nc * hc, ks1 * 64, device=dev)
    w1g[:h, :c] = w1 * gamma.detach().float()[None, :]
    w2p = torch.zeros(n2, nc * hc + 64, device=dev)
    if fold:
        w2p[:, :h] = (wa32.double() @ w2.double()).float()
    else:
        w2p[:c, :h] = w2
    bias1 = torch.zeros(nc * hc, device=dev)
    bias1[:h] = w1 @ beta.detach().float() + (fc1_b.detach().float() if fc1_b is not None else 0.0)
    colsum1 = w1g.to(torch.bfloat16).float().sum(dim=1)
    bias2 = torch.zeros(n2, device=dev)
    if fc2_b is not None and not fold:
        bias2[:c] = fc2_b.detach().float()
    slabs1, slabs2 = [], []
    for j, wj in enumerate(pl["widths"]):
        for s in range(ks1):
            slabs1.append(_swizzle_slab(w1g[j * hc:j * hc + wj, 64 * s:64 * s + 64].contiguous()))
        for s in range((wj + 63) // 64):
            k0 = j * hc + 64 * s
            valid = wj - 64 * s                      # hidden columns of this slab that belong to chunk j
            dcol = 0
            for rows in pl["pieces"]:
                blk = w2p[dcol:dcol + rows, k0:k0 + 64].clone()
                if valid < 64:
                    blk[:, valid:] = 0.0
                slabs2.append(_swizzle_slab(blk))
                dcol += rows
    pieces = pl["pieces"] + [0] * (2 - len(pl["pieces"]))
    plan = [ks1, pl["k1steps"], nc, hc, n2, pl["acc1_col"][0], pl["acc1_col"][1], len(pl["pieces"]), pieces[0], pieces[1],
            pl["w1_slots"], pl["w1_slot_bytes"], pl["w2_slots"], pl["w2_slot_bytes"]] + pl["widths"] + [0] * (8 - nc) + [pl["adj_tcol"], pl["fold"]]
    wadj = bias_adj = None
    if fuse_adj:
        wa = torch.zeros(_ADJ_N, ks1 * 64, device=dev)
        wa[:, :c] = wa32
        wadj = torch.cat([_swizzle_slab(wa[:, 64 * s:64 * s + 64].contiguous()) for s in range(ks1)]).contiguous()
        bias_adj = adjust_b.detach().float().clone() if adjust_b is not None else torch.zeros(_ADJ_N, device=dev)
        if fold and fc2_b is not None:
            bias_adj = bias_adj + (wa32.double() @ fc2_b.detach().double()).float()       # W_adj b2 joins the conv's own bias
        bias_adj = bias_adj.contiguous()
    return PackedMlp(torch.cat(slabs1).contiguous(), torch.cat(slabs2).contiguous(), bias1, colsum1, bias2,
                     torch.tensor(plan, dtype=torch.int32), float(eps), c, h, wadj, bias_adj)


def pack_swin_mlp_conv_res(fc1_w, fc1_b, gamma, beta, eps, fc2_w, fc2_b, conv_w, conv_b, alpha: float) -> PackedMlp:
    """norm2 + fc1 + GELU + fc2 of the RDG's LAST Swin block with adjust5 and the group's residual folded in
    (src/drct.py:394-396: `x5 = adjust5(swin5(...)); return x5 * 0.2 + x`) for adsr_swin_mlp_conv_res_bf16:
        out = res + alpha * (W_a z + b_a),  z = y + W2 g + b2   =>   out = res + y (alpha W_a)^T + g (alpha W_a W2)^T + alpha (b_a + W_a b2)
    The fc2 ring carries the c_out rows of alpha W_a W2 (with GELU's 0.5), led every tile by the (K slab, N piece) slabs of alpha W_a."""
    w1 = fc1_w.detach().float()
    h, c = w1.shape
    dev = w1.device
    wa = conv_w.detach().float().reshape(conv_w.shape[0], -1) * float(alpha)
    co = wa.shape[0]
    pl = swin_mlp_plan(c, h, False, co)
    hc, nc, n2, ks1 = pl["hc"], pl["nc"], pl["n2"], pl["ks1"]
    w1g = torch.zeros(nc * hc, ks1 * 64, device=dev)
    w1g[:h, :c] = w1 * gamma.detach().float()[None, :]
    w2p = torch.zeros(n2, nc * hc + 64, device=dev)
    w2p[:co, :h] = (wa.double() @ (fc2_w.detach().double() * 0.5)).float()
    wap = torch.zeros(n2, ks1 * 64, device=dev)
    wap[:co, :c] = wa
    bias1 = torch.zeros(nc * hc, device=dev)
    bias1[:h] = w1 @ beta.detach().float() + (fc1_b.detach().float() if fc1_b is not None else 0.0)
    colsum1 = w1g.to(torch.bfloat16).float().sum(dim=1)
    bias2 = torch.zeros(n2, device=dev)
    b = (conv_b.detach().double() * float(alpha)) if conv_b is not None else torch.zeros(co, device=dev, dtype=torch.float64)
    if fc2_b is not None:
        b = b + wa.double() @ fc2_b.detach().double()
    bias2[:co] = b.float()
    slabs1, slabs2 = [], []
    yslabs = []                                          # the y term: alpha W_a, (K slab, N piece)
    for s in range(ks1):
        dcol = 0
        for rows in pl["pieces"]:
            yslabs.append(_swizzle_slab(wap[dcol:dcol + rows, 64 * s:64 * s + 64].contiguous()))
            dcol += rows
    # ... ypre[j] of them lead chunk j's own slabs in the fc2 stream (the launcher derives the same schedule: all in front of chunk 0)
    ypre = [len(yslabs) if j == 0 else 0 for j in range(nc)]
    for j, wj in enumerate(pl["widths"]):
        for s in range(ks1):
            slabs1.append(_swizzle_slab(w1g[j * hc:j * hc + wj, 64 * s:64 * s + 64].contiguous()))
        for _ in range(ypre[j]):
            slabs2.append(yslabs.pop(0))
        for s in range((wj + 63) // 64):
            k0 = j * hc + 64 * s
            valid = wj - 64 * s
            dcol = 0
            for rows in pl["pieces"]:
                blk = w2p[dcol:dcol + rows, k0:k0 + 64].clone()
                if valid < 64:
                    blk[:, valid:] = 0.0
                slabs2.append(_swizzle_slab(blk))
                dcol += rows
    pieces = pl["pieces"] + [0] * (2 - len(pl["pieces"]))
    plan = [ks1, pl["k1steps"], nc, hc, n2, pl["acc1_col"][0], pl["acc1_col"][1], len(pl["pieces"]), pieces[0], pieces[1],
            pl["w1_slots"], pl["w1_slot_bytes"], pl["w2_slots"], pl["w2_slot_bytes"]] + pl["widths"] + [0] * (8 - nc) + [0, 2]
    pm = PackedMlp(torch.cat(slabs1).contiguous(), torch.cat(slabs2).contiguous(), bias1, colsum1, bias2,
                   torch.tensor(plan, dtype=torch.int32), float(eps), c, h)
    pm.conv_out = co
    return pm


# ------------------------------------------------------------------------------------------------ fused attention half
@dataclass
class PackedAttn:
    w1: torch.Tensor          # uint8: qkv slabs [3*hdp rows (q_h | k_h | v_h) x 64 bf16], (head, K slab) order, gamma folded
    w2: torch.Tensor          # uint8: proj slabs [cp rows x 64 bf16] (K = the head's channels), one per head
    bias_qkv: torch.Tensor    # fp32 [heads * 3 * hdp]
    colsum_qkv: torch.Tensor  # fp32 [heads * 3 * hdp]
    bias_p: torch.Tensor      # fp32 [cp]
    ln_eps: float
    C: int
    heads: int
    hd: int
    hdp: int


def pack_swin_attn(qkv_w, qkv_b, gamma, beta, eps, proj_w, proj_b, heads: int) -> PackedAttn:
    """norm1 + qkv Linear + proj Linear of one Swin block (src/drct.py:432, 245-249, 278, 300) for adsr_swin_attn_bf16:
    qkv rows regrouped per head as q_h | k_h | v_h (each padded to head_pad(hd) zero rows), gamma folded in
    (pack_ln_gemm_weight's algebra); proj columns split per head into [cp x 64] slabs."""
    w = qkv_w.detach().float()
    c = w.shape[1]
    dev = w.device
    hd = c // heads
    hdp = head_pad(hd)
    ks = (c + 63) // 64
    cp = round_up(c, 16)
    g, bt = gamma.detach().float(), beta.detach().float()
    wg = torch.zeros(heads, 3, hdp, ks * 64, device=dev)
    wg[:, :, :hd, :c] = (w * g[None, :]).view(3, heads, hd, c).permute(1, 0, 2, 3)
    t = w @ bt + (qkv_b.detach().float() if qkv_b is not None else 0.0)
    bias = torch.zeros(heads, 3, hdp, device=dev)
    bias[:, :, :hd] = t.view(3, heads, hd).permute(1, 0, 2)
    colsum = wg.to(torch.bfloat16).float().sum(dim=-1)
    slabs1 = []
    for h in range(heads):
        wh = wg[h].reshape(3 * hdp, ks * 64)
        for s in range(ks):
            slabs1.append(_swizzle_slab(wh[:, 64 * s:64 * s + 64].contiguous()))
    pw = proj_w.detach().float()
    slabs2 = []
    if hdp <= 64:
        for h in range(heads):
            blk = torch.zeros(cp, 64, device=dev)
            blk[:c, :hd] = pw[:, h * hd:(h + 1) * hd]
            slabs2.append(_swizzle_slab(blk))
    bias_p = torch.zeros(cp, device=dev)
    if proj_b is not None:
        bias_p[:c] = proj_b.detach().float()
    w2 = torch.cat(slabs2).contiguous() if slabs2 else torch.zeros(16, dtype=torch.uint8, device=dev)
    return PackedAttn(torch.cat(slabs1).contiguous(), w2, bias.reshape(-1).contiguous(), colsum.reshape(-1).contiguous(),
                      bias_p, float(eps), c, heads, hd, hdp)
