"""First slice of the training step (reference `Trainer.train` src/trainer.py:141-227, `Loss`
src/loss.py:83-84, 108-121, Adam src/trainer.py:49-59): the L1 loss with its gradient, the backward of the last layer
(`conv_last`) and a fused multi-tensor Adam, each one C-ABI call, plus the data-parallel gradient all-reduce (NCCL, one flat
bucket).  The backward of the tcgen05 blocks (attention, MLP, implicit-GEMM convs) does not exist yet, so `Trainer.train`
still raises; these pieces are parity-tested against torch autograd / torch.optim.Adam (tests/test_gpu_training.py)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi
from ._abi import check, lib, ptr, stream_ptr


def l1_loss_and_grad(sr: torch.Tensor, hr: torch.Tensor, grad_scale: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn.L1Loss()(sr, hr) (mean) and d loss / d sr * grad_scale, one pass over the two fp32 tensors.  -> (loss fp32 [1], grad like sr)."""
    if not sr.is_cuda:
        raise RuntimeError("sr must be a CUDA tensor: this package has no CPU fallback")
    sr, hr = sr.contiguous().float(), hr.contiguous().float()
    assert sr.shape == hr.shape
    n = sr.numel()
    grad = torch.empty_like(sr)
    nblk = max(1, min(148 * 8, (n // 4 + 255) // 256))
    ws = torch.empty(nblk, dtype=torch.float64, device=sr.device)
    loss = torch.empty(1, dtype=torch.float32, device=sr.device)
    check(lib().adsr_l1_loss_grad(ptr(sr), ptr(hr), n, float(grad_scale), ptr(grad), ptr(ws), nblk, ptr(loss), stream_ptr()), "adsr_l1_loss_grad")
    return loss, grad


def conv_last_backward(x: torch.Tensor, b: int, h: int, w: int, weight: torch.Tensor, grad_out: torch.Tensor, need_dx: bool = True):
    """Backward of conv_last (3x3, pad 1, 64 -> n_colors).  x: bf16 NHWC rows [b*h*w, >= 64] (the forward's input, e.g. the last
    PixelShuffle stage), weight fp32 [nc, 64, 3, 3], grad_out fp32 [b, nc, h, w] -> (dx bf16 [b*h*w, 64] or None, dw, db)."""
    if not x.is_cuda:
        raise RuntimeError("x must be a CUDA tensor: this package has no CPU fallback")
    nc, cin = weight.shape[0], weight.shape[1]
    grad_out, weight = grad_out.contiguous().float(), weight.detach().contiguous().float()
    dx = torch.empty(b * h * w, cin, dtype=torch.bfloat16, device=x.device) if need_dx else None
    dw, db = torch.empty_like(weight), torch.empty(nc, dtype=torch.float32, device=x.device)
    ws = torch.empty(lib().adsr_conv_last_bwd_workspace_bytes(b, h, cin) // 4, dtype=torch.float32, device=x.device)
    check(lib().adsr_conv_last_bwd(ptr(x), x.stride(0), ptr(grad_out), ptr(weight), b, h, w, cin, nc, ptr(dx), cin if need_dx else 8, ptr(dw),
                                   ptr(db), ptr(ws), stream_ptr()), "adsr_conv_last_bwd")
    return dx, dw, db


class FusedAdam:
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) (the reference's `make_optimizer`, src/trainer.py:49-59) as ONE kernel
    launch per step over all parameters: a device table of (param, grad, m, v, n) records and a chunk table (one block per 64 Ki
    elements).  Parameters and gradients are fp32; `step(grads)` takes the gradients in parameter order."""

    CHUNK = 65536

    def __init__(self, params: Sequence[torch.Tensor], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        self.params = [p for p in params]
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params):
            raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.step_count = 0
        chunks = [(i, c) for i, p in enumerate(self.params) for c in range((p.numel() + self.CHUNK - 1) // self.CHUNK)]
        self.n_chunks = len(chunks)
        dev = self.params[0].device
        self._chunks = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self._table = torch.empty(len(self.params), 5, dtype=torch.int64, device=dev)
        self._grad_ptrs: Optional[List[int]] = None

    def _fill_table(self, grads: Sequence[torch.Tensor]) -> None:
        ptrs = [g.data_ptr() for g in grads]
        if ptrs == self._grad_ptrs:
            return
        rows = [[p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()] for p, g, m, v in zip(self.params, grads, self.m, self.v)]
        self._table.copy_(torch.tensor(rows, dtype=torch.int64))
        self._grad_ptrs = ptrs

    @torch.no_grad()
    def step(self, grads: Sequence[torch.Tensor]) -> None:
        if len(grads) != len(self.params) or not all(g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.numel() == p.numel()
                                                       for g, p in zip(grads, self.params)):
            raise RuntimeError("FusedAdam.step needs one contiguous fp32 CUDA gradient per parameter")
        self._fill_table(grads)
        self.step_count += 1
        check(lib().adsr_adam_step(ptr(self._table), ptr(self._chunks), self.n_chunks, self.CHUNK, self.lr, self.betas[0], self.betas[1], self.eps,
                                   self.weight_decay, self.step_count, stream_ptr()), "adsr_adam_step")


def allreduce_gradients(grads: Sequence[torch.Tensor], average: bool = True) -> None:
    """Data-parallel gradient exchange (109.5 MB fp32 per DRCT-L step): ONE flat bucket, one all_reduce over
    NCCL / NVLink (gloo on CPU tests), copied back in place.  No-op without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.div_(dist.get_world_size())
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
