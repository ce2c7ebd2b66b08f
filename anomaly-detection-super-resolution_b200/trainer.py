"""The validation half of the reference's `src.trainer` (/root/reference/src/trainer.py): `quantize`, `calc_psnr`, `calc_ssim`
and `Trainer.test` (src/trainer.py:45-47, 95-99, 242-304), batched and on the GPU end to end.

  reference `Trainer.test`                                   here
  --------------------------------------------------------   -------------------------------------------------------------
  test loader yields batch 1                                  any batch size; per-image results are independent of it
  sr = model(lr); sr = quantize(sr, rgb_range)  (4 kernels)   rounding quantisation applied on the fly inside the metric kernel
  calc_psnr + calc_ssim: ~25 small kernels + 2 .item() syncs  ONE launch per batch (adsr_validate_images): PSNR + SSIM per image
  eval_psnr / len(loader_test), eval_ssim / len(loader_test)  same sums, in the loader's order

The training half (`Trainer.train`, loss, optimiser: SURVEY.md section 8f row 1) is outside the inference + scoring hot path:
`Trainer.train` raises NotImplementedError.  No CPU fallback.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Tuple

import torch

from . import metrics, ops


def quantize(img: torch.Tensor, rgb_range: float) -> torch.Tensor:
    """src/trainer.py:45-47 (kept for API parity; the validation loop below fuses it into the metric kernel)."""
    pixel_range = 255 / rgb_range
    return img.mul(pixel_range).clamp(0, 255).round().div(pixel_range)


def calc_psnr(sr, hr, scale, rgb_range, benchmark=False):
    return metrics.psnr_torch(sr, hr, rgb_range)


def calc_ssim(sr, hr, scale, rgb_range, benchmark=False):
    return metrics.ssim_torch(sr, hr, rgb_range, win_size=11)


def _shaved(t: torch.Tensor, shave: int = 4) -> torch.Tensor:
    return t[..., shave:-shave, shave:-shave] if t.size(-1) > 2 * shave else t     # src/metrics.py:73-75, 88-91


@torch.no_grad()
def validate_batch(model, lr: torch.Tensor, hr: torch.Tensor, rgb_range: float, win_size: int = 11) -> torch.Tensor:
    """One batch of the validation loop: fp64 device table [B, 2] = (PSNR, SSIM) per image of quantize(model(lr)) vs hr."""
    dev = torch.device("cuda")
    lr_d, hr_d = lr.to(dev, non_blocking=True).float(), hr.to(dev, non_blocking=True).float()
    target = model.model if hasattr(model, "model") else model
    sr = target(lr_d)
    if isinstance(sr, (list, tuple)):                       # DRN returns [x1, x2, x4]: the last one is evaluated (src/trainer.py:269)
        sr = sr[-1]
    if sr.size(-2) > hr_d.size(-2) or sr.size(-1) > hr_d.size(-1):                  # src/metrics.py:84-85
        sr = sr[..., :hr_d.size(-2), :hr_d.size(-1)]
    s = ops.validate_images(_shaved(sr), _shaved(hr_d), rgb_range, win_size)      # [B, 3] = ssim, mse, psnr
    return torch.stack([s[:, 2], s[:, 0]], dim=1)


@torch.no_grad()
def validate(model, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], rgb_range: float, win_size: int = 11) -> dict:
    """The body of `Trainer.test` over an iterable of (lr, hr) batches -> dict(psnr=[n], ssim=[n], eval_psnr, eval_ssim, n);
    eval_* are the loader averages the reference logs (an identical pair makes its PSNR, and the average, inf)."""
    rows: List[torch.Tensor] = [validate_batch(model, lr, hr, rgb_range, win_size) for lr, hr in batches]
    if not rows:
        return dict(psnr=[], ssim=[], eval_psnr=float("nan"), eval_ssim=float("nan"), n=0)
    table = torch.cat(rows).cpu().numpy()                   # the only device -> host read of the loop
    n = table.shape[0]
    return dict(psnr=table[:, 0], ssim=table[:, 1], eval_psnr=float(table[:, 0].sum() / n), eval_ssim=float(table[:, 1].sum() / n), n=n)


class Trainer:
    """`Trainer(opt, loader, my_model, my_loss, ckp)` with the reference's `test()`; training is not part of this build."""

    def __init__(self, opt, loader, my_model, my_loss=None, ckp=None, dual_model=False):
        self.opt, self.scale, self.ckp = opt, opt.scale, ckp
        self.loader_train = getattr(loader, "loader_train", None)
        self.loader_test = getattr(loader, "loader_test", loader)
        self.model, self.loss = my_model, my_loss
        self.best = {"psnr": (-math.inf, 0), "ssim": (-math.inf, 0)}
        self.epoch = 0

    def train(self):
        raise NotImplementedError("training is outside the B200 inference+scoring hot path (SURVEY.md 8f row 1)")

    def _pairs(self):
        for item in self.loader_test:                        # (lr, hr, filename) like src/data.py, or (lr, hr)
            lr, hr = item[0], item[1]
            if isinstance(lr, (list, tuple)):                # the reference's loader nests lr per scale: lr[0] is used
                lr = lr[0]
            if hr.nelement() == 1:                           # no_eval sample (src/trainer.py:258)
                continue
            yield lr, hr

    def test(self) -> Tuple[float, float]:
        """src/trainer.py:242-304: PSNR / SSIM of the round-quantised SR over the test loader, logged in the reference's format."""
        self.epoch += 1
        res = validate(self.model, self._pairs(), float(self.opt.rgb_range))
        n_loader = len(self.loader_test) if hasattr(self.loader_test, "__len__") else res["n"]
        denom = max(n_loader, 1)                             # the reference divides by len(loader_test), skipped samples included
        psnr, ssim = float(res["psnr"].sum() / denom) if res["n"] else 0.0, float(res["ssim"].sum() / denom) if res["n"] else 0.0
        for key, val in (("psnr", psnr), ("ssim", ssim)):
            if val > self.best[key][0]:
                self.best[key] = (val, self.epoch)
        s = max(self.scale) if isinstance(self.scale, (list, tuple)) else self.scale
        line = '[{} x{}]\tPSNR: {:.2f} (Best: {:.2f} @epoch {})\tSSIM: {:.4f} (Best: {:.4f} @epoch {})'.format(
            getattr(self.opt, "data_test", "test"), s, psnr, self.best["psnr"][0], self.best["psnr"][1], ssim, self.best["ssim"][0],
            self.best["ssim"][1])
        if self.ckp is not None and hasattr(self.ckp, "write_log"):
            self.ckp.write_log('\nEvaluation:')
            self.ckp.write_log(line)
        else:
            print(line)
        return psnr, ssim
