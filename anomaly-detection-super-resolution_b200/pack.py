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
    """(BN, n_tiles): BN multiple of 16, <= 256, minimal padding."""
    n16 = round_up(n, 16)
    n_tiles = (n16 + 255) // 256
    bn = round_up((n16 + n_tiles - 1) // n_tiles, 16)
    return bn, n_tiles


@dataclass
class PackedWeight:
    data: torch.Tensor        # uint8 [n_tiles * k_stages * BN * 128]
    bias: torch.Tensor        # fp32 [n_tiles * BN]
    N: int                    # logical output columns
    K: int                    # logical K (GEMM) or Cin (conv)
    BN: int
    n_tiles: int
    k_stages: int


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
    return PackedWeight(data, _pad_bias(bias, n, bn * n_tiles, w.device), n, k, bn, n_tiles, ks)


def pack_conv3x3_weight(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> PackedWeight:
    """weight: [Cout, Cin, 3, 3] -> K ordered [tap = ky*3+kx][cin padded to a multiple of 64]."""
    w = weight.detach().float()
    cout, cin = w.shape[0], w.shape[1]
    spt = (cin + 63) // 64
    bn, n_tiles = choose_bn(cout)
    w2 = torch.zeros(cout, 9, spt * 64, dtype=torch.float32, device=w.device)
    w2[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    data, ks = _swizzle_tiles(w2.reshape(cout, 9 * spt * 64), bn, n_tiles)
    return PackedWeight(data, _pad_bias(bias, cout, bn * n_tiles, w.device), cout, cin, bn, n_tiles, ks)


def head_pad(hd: int) -> int:
    return round_up(hd, 16)


def pack_qkv_weight(weight: torch.Tensor, bias: Optional[torch.Tensor], heads: int) -> PackedWeight:
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
