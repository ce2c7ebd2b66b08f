"""Drop-in for the reference's `src.model` (/root/reference/src/model.py:46-170): `make_model(opt)` and the
thin `Model(opt, ckp)` wrapper with the same forward / save / load contract (bare state_dict files,
`strict=False` loading, `get_model`, `state_dict`).  The device is always CUDA: there is no CPU path.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .drct import DRCT


def make_model(opt):
    if opt.model_name == 'drct':
        return DRCT(opt)
    elif opt.model_name == 'drn-l':
        from .drn import DRN

        return DRN(opt)
    else:
        print(f"No model with this name: {opt.model_name}")     # same behaviour as src/model.py:51-52


class Model(nn.Module):
    def __init__(self, opt, ckp=None, dual_model=False):
        super().__init__()
        print('Making model...')
        self.opt = opt
        self.scale = opt.scale
        self.idx_scale = 0
        self.self_ensemble = getattr(opt, 'self_ensemble', False)
        if getattr(opt, 'cpu', False) or not torch.cuda.is_available():
            raise RuntimeError("the B200 build runs on CUDA only (opt.cpu / missing GPU): no CPU fallback")
        self.cpu = False
        self.device = torch.device('cuda')
        if ckp is not None:
            ckp.write_log(f"Using device: {self.device}")
        self.n_GPUs = getattr(opt, 'n_GPUs', 1)
        self.dual_model = False            # the dual (training-only) DownBlocks are not part of the inference path
        self.model = make_model(opt).to(self.device).eval()
        self.load(getattr(opt, 'pre_train', '.'), getattr(opt, 'pre_train_dual', '.'))
        num_parameter = self.count_parameters(self.model)
        if ckp is not None:
            ckp.write_log(f"The number of parameters is {num_parameter / 1000 ** 2:.2f}M")

    def forward(self, x, idx_scale=0):
        self.idx_scale = idx_scale
        return self.model(x)

    def get_model(self):
        return self.model

    def state_dict(self, **kwargs):
        return self.get_model().state_dict(**kwargs)

    def count_parameters(self, model):
        return sum(p.numel() for p in model.parameters() if p.requires_grad)

    def save(self, path, is_best=False):
        target = self.get_model()
        os.makedirs(os.path.join(path, 'model'), exist_ok=True)
        torch.save(target.state_dict(), os.path.join(path, 'model', 'model_latest.pt'))
        if is_best:
            torch.save(target.state_dict(), os.path.join(path, 'model', 'model_best.pt'))

    def load(self, pre_train='.', pre_train_dual='.', cpu=False):
        if pre_train != '.':
            print('Loading model from {}'.format(pre_train))
            sd = torch.load(pre_train, weights_only=True, map_location=self.device)
            self.get_model().load_state_dict(sd, strict=False)
