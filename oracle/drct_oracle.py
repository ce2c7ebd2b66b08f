"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference DRCT forward pass.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import this file.  The product package never does (it fails loudly without its CUDA library).

What is restated (citations are into /root/reference):
  * DRCT.forward / forward_features ............ src/drct.py:870-899
  * RDG.forward + block table .................. src/drct.py:322-396
  * SwinTransformerBlock.forward ............... src/drct.py:472-512
  * torch.roll + window_partition/reverse ...... src/drct.py:193-220, 482-505 (as closed-form index maps)
  * calculate_mask ............................. src/drct.py:449-470 (as closed-form region ids)
  * WindowAttention.forward .................... src/drct.py:271-302
  * relative_position_index .................... src/drct.py:246-257
  * Mlp.forward ................................ src/drct.py:173-190
  * PatchEmbed/UnEmbed, Upsample ............... src/drct.py:621-713

All arithmetic lives in torch (third party, torch>=2.0, lock torch==2.8.0 in the reference's
requirements.lock.txt:60; this image has 2.11 with identical semantics for every op used).
The reference's own tests hold no golden vectors for this path, so the oracle is
pinned against the reference code itself, executed in the build container: see
`oracle/make_golden.py` (writes tests/golden/*.npz) and tests/test_oracle_pinning.py.
Evaluation mode is assumed (DropPath = identity), see SURVEY.md section 0 item 2.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

RGB_MEAN = (0.4488, 0.4371, 0.4040)  # src/drct.py:774


@dataclass
class DrctCfg:
    img_size: int = 32          # LR side (opt.img_size)
    n_colors: int = 3
    embed_dim: int = 180
    num_layers: int = 12        # len(opt.depths)
    num_heads: int = 6          # opt.num_heads[i] (all equal in the reference configs)
    window_size: int = 8        # opt.window_size = img_size // 4  (src/main.py:286)
    mlp_ratio: float = 2.0
    upscale: int = 4
    img_range: float = 1.0
    gc: int = 32
    num_feat: int = 64          # src/drct.py:771

    def block_table(self) -> List[dict]:
        """Per-RDG Swin block table (src/drct.py:326-374)."""
        d, g, nh, ws = self.embed_dim, self.gc, self.num_heads, self.window_size
        out = []
        for k in range(5):
            dim = d + k * g
            heads = nh if k == 0 else nh - (dim % nh)
            ratio = self.mlp_ratio if k < 3 else 1
            out.append(dict(dim=dim, heads=heads, head_dim=dim // heads,
                            shift=(ws // 2) if k in (1, 3) else 0,
                            hidden=int(dim * ratio),
                            adjust_out=g if k < 4 else d))
        return out


# --------------------------------------------------------------------------------------------
# index maps (bit-exact objects)
# --------------------------------------------------------------------------------------------
def window_source_index(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """int64 [nW, ws*ws]: flat pixel index (y*W+x) that lands in window w, slot n after
    roll(-shift,-shift) followed by window_partition (src/drct.py:482-489, 193-204)."""
    nwx = W // ws
    w = torch.arange((H // ws) * nwx).view(-1, 1)
    n = torch.arange(ws * ws).view(1, -1)
    ys = (w // nwx) * ws + n // ws          # coordinates in the shifted frame
    xs = (w % nwx) * ws + n % ws
    return ((ys + shift) % H) * W + (xs + shift) % W


def shift_region_ids(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """int64 [nW, ws*ws] region counter `cnt` of calculate_mask (src/drct.py:452-463) per window slot."""
    nwx = W // ws
    w = torch.arange((H // ws) * nwx).view(-1, 1)
    n = torch.arange(ws * ws).view(1, -1)
    ys = (w // nwx) * ws + n // ws
    xs = (w % nwx) * ws + n % ws

    def r(t, L):
        return (t >= L - ws).long() + (t >= L - shift).long()

    return 3 * r(ys, H) + r(xs, W)


def attention_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """fp32 [nW, N, N]: 0 where region ids agree, -100.0 elsewhere (src/drct.py:465-468).
    mask[w, q, k] = cnt[k] - cnt[q] != 0 (mask_windows.unsqueeze(1) - mask_windows.unsqueeze(2))."""
    ids = shift_region_ids(H, W, ws, shift)
    neq = ids[:, None, :] != ids[:, :, None]
    return torch.where(neq, torch.tensor(-100.0), torch.tensor(0.0))


def relative_position_index(ws: int) -> torch.Tensor:
    """int64 [N, N] (src/drct.py:246-257): (yq-yk+ws-1)*(2ws-1) + (xq-xk+ws-1)."""
    n = torch.arange(ws * ws)
    y, x = n // ws, n % ws
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


# --------------------------------------------------------------------------------------------
# deterministic synthetic weights in the reference's state_dict layout
# --------------------------------------------------------------------------------------------
def _trunc_normal(gen, shape, std):
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0, generator=gen)
    return t


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def make_state_dict(cfg: DrctCfg, seed: int = 1, affine_jitter: float = 0.0) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's key/shape layout (SURVEY.md section 8b 'State dict').

    Init follows the reference's rules in spirit (trunc-normal 0.02 linears, src/drct.py:851-858;
    PyTorch default uniform convs) but uses its own generator so it is reproducible on any box.
    `affine_jitter` > 0 perturbs LayerNorm affine and all biases so that tests exercise them.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    ws, N = cfg.window_size, cfg.window_size ** 2

    def conv(name, cout, cin, k):
        bound = 1.0 / math.sqrt(cin * k * k)
        sd[f"{name}.weight"] = _uniform(g, (cout, cin, k, k), bound)
        sd[f"{name}.bias"] = _uniform(g, (cout,), bound)

    def linear(name, cout, cin):
        sd[f"{name}.weight"] = _trunc_normal(g, (cout, cin), 0.02)
        sd[f"{name}.bias"] = (_uniform(g, (cout,), affine_jitter) if affine_jitter else torch.zeros(cout))

    def norm(name, c):
        sd[f"{name}.weight"] = torch.ones(c) + (_uniform(g, (c,), affine_jitter) if affine_jitter else 0)
        sd[f"{name}.bias"] = (_uniform(g, (c,), affine_jitter) if affine_jitter else torch.zeros(c))

    conv("conv_first", cfg.embed_dim, cfg.n_colors, 3)
    norm("patch_embed.norm", cfg.embed_dim)
    for i in range(cfg.num_layers):
        for k, blk in enumerate(cfg.block_table(), start=1):
            p = f"layers.{i}.swin{k}"
            dim, nh = blk["dim"], blk["heads"]
            if blk["shift"] > 0:
                sd[f"{p}.attn_mask"] = attention_mask(cfg.img_size, cfg.img_size, ws, blk["shift"])
            norm(f"{p}.norm1", dim)
            sd[f"{p}.attn.relative_position_bias_table"] = _trunc_normal(g, ((2 * ws - 1) ** 2, nh), 0.02)
            sd[f"{p}.attn.relative_position_index"] = relative_position_index(ws)
            linear(f"{p}.attn.qkv", 3 * dim, dim)
            linear(f"{p}.attn.proj", dim, dim)
            norm(f"{p}.norm2", dim)
            linear(f"{p}.mlp.fc1", blk["hidden"], dim)
            linear(f"{p}.mlp.fc2", dim, blk["hidden"])
            conv(f"layers.{i}.adjust{k}", blk["adjust_out"], dim, 1)
    norm("norm", cfg.embed_dim)
    conv("conv_after_body", cfg.embed_dim, cfg.embed_dim, 3)
    conv("conv_before_upsample.0", cfg.num_feat, cfg.embed_dim, 3)
    for j in range(int(math.log2(cfg.upscale))):
        conv(f"upsample.{2 * j}", 4 * cfg.num_feat, cfg.num_feat, 3)
    conv("conv_last", cfg.n_colors, cfg.num_feat, 3)
    return sd


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    acc = 0.0
    for i, k in enumerate(sorted(sd)):
        acc += float(sd[k].double().abs().sum()) * (1 + (i % 7))
    return acc


# --------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------
def window_attention(xw: torch.Tensor, sd, p: str, heads: int, ws: int, mask) -> torch.Tensor:
    """src/drct.py:271-302.  xw: [B_, N, C] windows."""
    B_, N, C = xw.shape
    hd = C // heads
    qkv = F.linear(xw, sd[f"{p}.qkv.weight"], sd[f"{p}.qkv.bias"])
    qkv = qkv.view(B_, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    table = sd[f"{p}.relative_position_bias_table"]
    bias = table[relative_position_index(ws).reshape(-1)].view(N, N, heads).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask[None, :, None]).view(B_, heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(out, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])


def swin_block(x: torch.Tensor, H: int, W: int, sd, p: str, blk: dict, ws: int, taps=None) -> torch.Tensor:
    """src/drct.py:472-512.  x: [B, L, C]."""
    B, L, C = x.shape
    shift = blk["shift"]
    h = F.layer_norm(x, (C,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-5)
    src = window_source_index(H, W, ws, shift)                     # [nW, N]
    nW, N = src.shape
    xw = h[:, src.reshape(-1), :].view(B * nW, N, C)               # roll + partition as one gather
    mask = attention_mask(H, W, ws, shift) if shift > 0 else None
    aw = window_attention(xw, sd, f"{p}.attn", blk["heads"], ws, mask)
    a = torch.empty_like(x)
    a[:, src.reshape(-1), :] = aw.view(B, nW * N, C)               # reverse + roll back = inverse scatter
    x = x + a
    if taps is not None:
        taps[f"{p}.attn_res"] = x
    h = F.layer_norm(x, (C,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-5)
    h = F.linear(h, sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"])
    h = F.gelu(h)                                                  # exact erf (nn.GELU default)
    h = F.linear(h, sd[f"{p}.mlp.fc2.weight"], sd[f"{p}.mlp.fc2.bias"])
    return x + h


def rdg(x: torch.Tensor, H: int, W: int, sd, i: int, cfg: DrctCfg, taps=None) -> torch.Tensor:
    """src/drct.py:388-396 on tokens [B, L, C] (PatchEmbed/UnEmbed are pure transposes)."""
    feats = [x]
    table = cfg.block_table()
    for k, blk in enumerate(table, start=1):
        inp = torch.cat(feats, dim=-1)
        y = swin_block(inp, H, W, sd, f"layers.{i}.swin{k}", blk, cfg.window_size, taps)
        w = sd[f"layers.{i}.adjust{k}.weight"]
        y = F.linear(y, w.view(w.shape[0], w.shape[1]), sd[f"layers.{i}.adjust{k}.bias"])  # 1x1 conv
        if k < 5:
            y = F.leaky_relu(y, 0.2)
        if taps is not None:
            taps[f"layers.{i}.x{k}"] = y
        feats.append(y)
    return feats[-1] * 0.2 + x


def drct_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: DrctCfg, taps=None) -> torch.Tensor:
    """x: [B, nc, h, w] fp32 in [0, rgb_range]  ->  [B, nc, h*s, w*s]  (src/drct.py:886-899)."""
    x = x.float()
    B, nc, H, W = x.shape
    mean = torch.tensor(RGB_MEAN).view(1, 3, 1, 1) if nc == 3 else torch.zeros(1, 1, 1, 1)
    x = (x - mean) * cfg.img_range
    f0 = F.conv2d(x, sd["conv_first.weight"], sd["conv_first.bias"], padding=1)
    t = f0.flatten(2).transpose(1, 2)                                          # [B, L, C]
    t = F.layer_norm(t, (cfg.embed_dim,), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"], 1e-5)
    if taps is not None:
        taps["embed"] = t
    for i in range(cfg.num_layers):
        t = rdg(t, H, W, sd, i, cfg, taps)
        if taps is not None:
            taps[f"layers.{i}.out"] = t
    t = F.layer_norm(t, (cfg.embed_dim,), sd["norm.weight"], sd["norm.bias"], 1e-5)
    f = t.transpose(1, 2).reshape(B, cfg.embed_dim, H, W)
    f = F.conv2d(f, sd["conv_after_body.weight"], sd["conv_after_body.bias"], padding=1) + f0
    f = F.leaky_relu(F.conv2d(f, sd["conv_before_upsample.0.weight"], sd["conv_before_upsample.0.bias"], padding=1), 0.01)
    for j in range(int(math.log2(cfg.upscale))):
        f = F.pixel_shuffle(F.conv2d(f, sd[f"upsample.{2 * j}.weight"], sd[f"upsample.{2 * j}.bias"], padding=1), 2)
    y = F.conv2d(f, sd["conv_last.weight"], sd["conv_last.bias"], padding=1)
    return y / cfg.img_range + mean


def flops_per_image(cfg: DrctCfg, h: int | None = None) -> float:
    """Algorithmic FLOPs (2*MAC, matmul+conv only, unpadded) per image -- SURVEY.md section 8d."""
    h = h or cfg.img_size
    L = h * h
    N = cfg.window_size ** 2
    tot = 0.0
    for blk in cfg.block_table():
        C, Hd = blk["dim"], blk["hidden"]
        tot += 2 * L * (4 * C * C + 2 * C * Hd)        # qkv + proj + fc1 + fc2
        tot += 4 * L * N * C                           # QK^T and PV
        tot += 2 * L * C * blk["adjust_out"]           # 1x1 adjust conv
    tot *= cfg.num_layers
    d, nf, nc = cfg.embed_dim, cfg.num_feat, cfg.n_colors
    tot += 2 * L * 9 * (nc * d + d * d + d * nf)
    px = L
    for _ in range(int(math.log2(cfg.upscale))):
        tot += 2 * px * 9 * nf * 4 * nf
        px *= 4
    tot += 2 * px * 9 * nf * nc
    return tot
