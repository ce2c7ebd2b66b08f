"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference DRN (dual regression network) forward.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this file.

Restated (citations into /root/reference):
  * DRN.__init__ channel plan / DRN.forward ....... src/drn.py:160-270
  * RCAB / CALayer ................................. src/drn.py:123-158 (res_scale is stored but never applied)
  * DownBlock ...................................... src/drn.py:83-119 (stride-2 conv + LeakyReLU(negval), conv; no bias)
  * Upsampler (conv + PixelShuffle(2)) ............. src/drn.py:55-81
  * MeanShift sub_mean / add_mean .................. src/drn.py:44-52, 176-185
  * nn.Upsample(bicubic, align_corners=False) ...... src/drn.py:174-175
Pinned against the reference code itself by oracle/make_golden.py -> tests/golden/drn_*.npz.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List

import torch
import torch.nn.functional as F


@dataclass
class DrnCfg:
    scale: int = 4               # total upscale; opt.scale = [2, 4] (src/main.py:169)
    n_blocks: int = 40
    n_feats: int = 20
    n_colors: int = 3
    rgb_range: float = 255.0
    negval: float = 0.2

    @property
    def phase(self) -> int:
        return int(math.log2(self.scale))


def make_state_dict(cfg: DrnCfg, seed: int = 1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    nc, nf, ph = cfg.n_colors, cfg.n_feats, cfg.phase

    def conv(name, cout, cin, k, bias=True, gain=1.0):
        bound = gain / math.sqrt(cin * k * k)
        sd[f"{name}.weight"] = (torch.rand((cout, cin, k, k), generator=g) * 2 - 1) * bound
        if bias:
            sd[f"{name}.bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound

    mean = torch.tensor((0.4488, 0.4371, 0.4040) if nc == 3 else (0.4440,))
    for name, sign in (("sub_mean", -1.0), ("add_mean", 1.0)):
        sd[f"{name}.weight"] = torch.eye(nc).view(nc, nc, 1, 1)
        sd[f"{name}.bias"] = sign * cfg.rgb_range * mean
    conv("head", nf, nc, 3)
    for p in range(ph):
        c = nf * 2 ** p
        conv(f"down.{p}.dual_module.0.0", c, c, 3, bias=False)
        conv(f"down.{p}.dual_module.1", 2 * c, c, 3, bias=False)
    for idx in range(ph):
        c = nf * 2 ** ph if idx == 0 else 2 * nf * 2 ** (ph - idx)
        for j in range(cfg.n_blocks):
            b = f"up_blocks.{idx}.{j}.body"
            conv(f"{b}.0", c, c, 3, gain=0.5)          # keep activations bounded over 40 residual blocks
            conv(f"{b}.2", c, c, 3, gain=0.5)
            conv(f"{b}.3.conv_du.0", c // 16, c, 1)
            conv(f"{b}.3.conv_du.2", c, c // 16, 1)
        conv(f"up_blocks.{idx}.{cfg.n_blocks}.0", 4 * c, c, 3)
        conv(f"up_blocks.{idx}.{cfg.n_blocks + 1}", nf * 2 ** (ph - idx - 1), c, 1)
    conv("tail.0", nc, nf * 2 ** ph, 3)
    for i, p in enumerate(range(ph, 0, -1), start=1):
        conv(f"tail.{i}", nc, nf * 2 ** p, 3)
    return sd


def state_dict_checksum(sd) -> float:
    return float(sum(float(v.double().abs().sum()) * (1 + (i % 7)) for i, (k, v) in enumerate(sorted(sd.items()))))


def rcab(x, sd, p):
    r = F.relu(F.conv2d(x, sd[f"{p}.body.0.weight"], sd[f"{p}.body.0.bias"], padding=1))
    r = F.conv2d(r, sd[f"{p}.body.2.weight"], sd[f"{p}.body.2.bias"], padding=1)
    y = r.mean(dim=(2, 3), keepdim=True)
    y = F.relu(F.conv2d(y, sd[f"{p}.body.3.conv_du.0.weight"], sd[f"{p}.body.3.conv_du.0.bias"]))
    y = torch.sigmoid(F.conv2d(y, sd[f"{p}.body.3.conv_du.2.weight"], sd[f"{p}.body.3.conv_du.2.bias"]))
    return r * y + x


def drn_forward(sd, x: torch.Tensor, cfg: DrnCfg) -> List[torch.Tensor]:
    """x: [B, nc, h, w] in [0, rgb_range] -> [sr_x1, sr_x2, ..., sr_xS] (src/drn.py:241-270)."""
    ph = cfg.phase
    x = F.interpolate(x.float(), scale_factor=cfg.scale, mode="bicubic", align_corners=False)
    x = F.conv2d(x, sd["sub_mean.weight"], sd["sub_mean.bias"])
    x = F.conv2d(x, sd["head.weight"], sd["head.bias"], padding=1)
    copies = []
    for p in range(ph):
        copies.append(x)
        x = F.leaky_relu(F.conv2d(x, sd[f"down.{p}.dual_module.0.0.weight"], None, stride=2, padding=1), cfg.negval)
        x = F.conv2d(x, sd[f"down.{p}.dual_module.1.weight"], None, padding=1)
    add_mean = lambda t: F.conv2d(t, sd["add_mean.weight"], sd["add_mean.bias"])
    results = [add_mean(F.conv2d(x, sd["tail.0.weight"], sd["tail.0.bias"], padding=1))]
    for idx in range(ph):
        for j in range(cfg.n_blocks):
            x = rcab(x, sd, f"up_blocks.{idx}.{j}")
        u = f"up_blocks.{idx}.{cfg.n_blocks}.0"
        x = F.pixel_shuffle(F.conv2d(x, sd[f"{u}.weight"], sd[f"{u}.bias"], padding=1), 2)
        c1 = f"up_blocks.{idx}.{cfg.n_blocks + 1}"
        x = F.conv2d(x, sd[f"{c1}.weight"], sd[f"{c1}.bias"])
        x = torch.cat((x, copies[ph - idx - 1]), 1)
        results.append(add_mean(F.conv2d(x, sd[f"tail.{idx + 1}.weight"], sd[f"tail.{idx + 1}.bias"], padding=1)))
    return results


def flops_per_image(cfg: DrnCfg, h: int) -> float:
    """Algorithmic conv FLOPs (2*MAC) per image for an h x h LR input (49.877 G for DRN-L x4, h=32)."""
    nc, nf, ph = cfg.n_colors, cfg.n_feats, cfg.phase
    H = h * cfg.scale
    px = lambda s: (H // s) ** 2
    tot = 2 * px(1) * (nc * nc + 9 * nc * nf)                  # sub_mean + head
    for p in range(ph):
        c = nf * 2 ** p
        tot += 2 * px(2 ** (p + 1)) * 9 * (c * c + c * 2 * c)
    tot += 2 * px(2 ** ph) * (9 * nf * 2 ** ph * nc + nc * nc)
    for idx in range(ph):
        c = nf * 2 ** ph if idx == 0 else 2 * nf * 2 ** (ph - idx)
        s = 2 ** (ph - idx)
        tot += cfg.n_blocks * (2 * px(s) * 9 * 2 * c * c + 2 * 2 * c * (c // 16))
        tot += 2 * px(s) * 9 * c * 4 * c
        cout = nf * 2 ** (ph - idx - 1)
        tot += 2 * px(s // 2) * c * cout
        tot += 2 * px(s // 2) * (9 * (cout * 2) * nc + nc * nc)
    return float(tot)
