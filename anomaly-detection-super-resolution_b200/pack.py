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
from dataclasses import dataclass
from typing import Optional

import torch


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


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


def _swizzle_tiles(w2d: torch.Tensor, bn: int, n_tiles: int) -> tuple[torch.Tensor, int]:
    """w2d: fp32 [N, Kpad] with Kpad % 64 == 0 -> packed uint8 image."""
    n, kpad = w2d.shape
    ks = kpad // 64
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


def pack_gemm_weight(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> PackedWeight:
    """weight: [N, K] (nn.Linear / 1x1 conv layout)."""
    w = weight.detach().float().reshape(weight.shape[0], -1)
    n, k = w.shape
    bn, n_tiles = choose_bn(n)
    kpad = round_up(k, 64)
    w2 = torch.zeros(n, kpad, dtype=torch.float32, device=w.device)
    w2[:, :k] = w
    data, ks = _swizzle_tiles(w2, bn, n_tiles)
    return PackedWeight(data, _pad_bias(bias, n, bn * n_tiles, w.device), N=n, K=k, BN=bn, n_tiles=n_tiles, k_stages=ks)


def pack_ln_gemm_weight(weight: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                        eps: float = 1e-5) -> PackedWeight:
    """Linear applied to LayerNorm(x) with the normalisation folded into the GEMM:
         LN(x) W^T + b = rstd * (x (gamma*W)^T - mean * s) + t,   s_n = sum_k gamma_k W_nk,   t_n = sum_k beta_k W_nk + b_n.
    The kernel multiplies RAW rows with the packed gamma*W and applies mean / rstd / s / t in its epilogue; s is taken from
    the bf16-rounded packed weights so that the subtraction cancels exactly what the tensor core accumulated."""
    w = weight.detach().float().reshape(weight.shape[0], -1)
    wg = w * gamma.detach().float()[None, :]
    t = w @ beta.detach().float() + (bias.detach().float() if bias is not None else 0.0)
    pw = pack_gemm_weight(wg, t)
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
    return PackedWeight(data, _pad_bias(bias, cout, bn * n_tiles, w.device), N=cout, K=cin, BN=bn, n_tiles=n_tiles, k_stages=ks)


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
        return pack_ln_gemm_weight(w3.view(3 * heads * hdp, c), b3.view(-1), gamma, beta, eps)
    return pack_gemm_weight(w3.view(3 * heads * hdp, c), b3.view(-1))


def pack_proj_weight(weight: torch.Tensor, bias: Optional[torch.Tensor], heads: int) -> PackedWeight:
    """proj Linear [C, C]: input columns regrouped to the head-padded attention output layout."""
    w = weight.detach().float()
    c = w.shape[1]
    hd = c // heads
    hdp = head_pad(hd)
    w2 = torch.zeros(w.shape[0], heads, hdp, dtype=torch.float32, device=w.device)
    w2[:, :, :hd] = w.view(w.shape[0], heads, hd)
    return pack_gemm_weight(w2.view(w.shape[0], heads * hdp), bias)
