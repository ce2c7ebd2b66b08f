"""Drop-in for the configuration surface of the reference's `src/main.py`: the `DRN` / `DRCT` option
dataclasses (src/main.py:35-142) and `setup_opt_drn` / `setup_opt_drct` (src/main.py:144-205, 243-294),
which are the only source of hyper-parameters the model constructors read.  Field names, defaults and the
positional signatures are the reference's; the training loop itself (Trainer / Loss / Checkpoint) is outside
this hot path ("next" row 1) and is not provided.
"""
from __future__ import annotations

import argparse
import os
import sys
from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class DRN:
    model_name: str = 'drn-l'
    n_threads: int = -2
    cpu: bool = False
    n_GPUs: int = 1
    seed: int = 1
    data_dir: str = './workspace/gkd/DC2/unlabeled/HR_512_grayscale/'
    data_train: str = ''
    data_test: str = ''
    data_range: str = '1-224/225-280'
    scale: int | list[int] = 4
    patch_size: int = 512
    rgb_range: int = 255
    n_colors: int = 1
    no_augment: bool = False
    pre_train: str = '.'
    pre_train_dual: str = '.'
    n_blocks: int = 40
    n_feats: int = 20
    negval: float = 0.2
    test_every: int = 10
    epochs: int = 10
    batch_size: int = 4
    self_ensemble: bool = False
    test_only: bool = False
    lr: float = 1e-4
    eta_min: float = 1e-7
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-8
    weight_decay: float = 1e-8
    loss: str = '1*L1'
    skip_threshold: float = 1.5
    dual_weight: float = 0.1
    save: str = './workspace/experiment/drn-l/gkd_dc2_unlabeld_X4_10_grayscale/'
    print_every: int = 10
    save_results: bool = True
    dual: bool = True
    patience: int = 10
    min_delta: float = 0.0
    dataset: str = ''
    classe: str = ''
    slurm: bool = False
    ssim_window_size: int = 11
    best_auc: float = 1.0


@dataclass
class DRCT:
    model_name: str = 'drct'
    n_threads: int = 1
    cpu: bool = False
    n_GPUs: int = 1
    seed: int = 1
    data_dir: str = './workspace/gkd/DC2/unlabeled/HR_512_grayscale/'
    data_train: str = ''
    data_test: str = ''
    data_range: str = '1-260/261-299'
    scale: int | list[int] = 4
    patch_size: int = 512
    rgb_range: int = 255
    n_colors: int = 1
    no_augment: bool = False
    pre_train: str = '.'
    pre_train_dual: str = '.'
    negval: float = 0.2
    test_every: int = 30
    epochs: int = 10
    batch_size: int = 2
    self_ensemble: bool = False
    test_only: bool = False
    lr: float = 1e-4
    eta_min: float = 1e-7
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-8
    loss: str = '1*L1'
    skip_threshold: float = 1e6
    dual_weight: float = 0.1
    save: str = './workspace/experiment/drct/gkd_dc2_unlabeled_X4_10_test_grayscale/'
    print_every: int = 10
    save_results: bool = True
    dual: bool = False
    upscale: int = 4
    img_size: int = 128
    window_size: int = 16
    compress_ratio: int = 3
    squeeze_factor: int = 30
    conv_scale: float = 0.01
    overlap_ratio: float = 0.5
    img_range: float = 1.0
    depths: tuple[int, ...] = (6,) * 12
    embed_dim: int = 180
    num_heads: tuple[int, ...] = (6,) * 12
    mlp_ratio: int = 2
    upsampler: str = 'pixelshuffle'
    resi_connection: str = '1conv'
    ema_decay: float = 0.999
    weight_decay: float = 0.0
    betas: tuple[float, float] = (0.9, 0.99)
    patience: int = 10
    min_delta: float = 0.0
    dataset: str = ''
    classe: str = ''
    slurm: bool = False
    ssim_window_size: int = 11
    best_auc: float = 1.0


def setup_opt_drn(opt: DRN, best_auc, ssim_window_size, dataset, classe, slurm, scale, no_augment, n_colors, epochs,
                  batch_size, patch_size, data_dir, save, data_range, test_every, print_every, patience, min_delta,
                  n_threads, pre_trained, pre_trained_dual, loss) -> DRN:
    """Positional signature and effects of src/main.py:144-205 (note: data_range is accepted but not stored)."""
    opt.scale = [pow(2, s + 1) for s in range(int(np.log2(scale)))]
    if scale == 2:
        opt.n_blocks, opt.n_feats = 44, 40
    elif scale == 4:
        opt.n_blocks, opt.n_feats = 40, 20
    elif scale == 8:
        opt.n_blocks, opt.n_feats = 36, 10
    else:
        print(f"No setup for this scale: {scale}")
    opt.no_augment, opt.n_colors, opt.epochs, opt.batch_size, opt.patch_size = no_augment, n_colors, epochs, batch_size, patch_size
    opt.data_dir, opt.save, opt.test_every, opt.print_every = data_dir, save, test_every, print_every
    opt.patience, opt.min_delta, opt.n_threads = patience, min_delta, n_threads
    opt.pre_train, opt.pre_train_dual, opt.loss = pre_trained, pre_trained_dual, loss
    opt.dataset, opt.classe, opt.slurm, opt.ssim_window_size, opt.best_auc = dataset, classe, slurm, ssim_window_size, best_auc
    return opt


def setup_opt_drct(opt: DRCT, best_auc, ssim_window_size, dataset, classe, slurm, scale, no_augment, n_colors, epochs,
                   batch_size, patch_size, img_size, data_dir, save, data_range, test_every, print_every, patience,
                   min_delta, n_threads, pre_trained, loss) -> DRCT:
    """Positional signature and effects of src/main.py:243-294 (window_size = img_size // 4, :286)."""
    opt.upscale = scale
    opt.scale = [scale]
    opt.no_augment, opt.n_colors, opt.epochs, opt.batch_size, opt.patch_size = no_augment, n_colors, epochs, batch_size, patch_size
    opt.data_dir, opt.data_range, opt.save = data_dir, data_range, save
    opt.test_every, opt.print_every, opt.img_size = test_every, print_every, img_size
    opt.patience, opt.min_delta, opt.n_threads, opt.pre_train = patience, min_delta, n_threads, pre_trained
    opt.window_size = img_size // 4
    opt.loss, opt.dataset, opt.classe, opt.slurm = loss, dataset, classe, slurm
    opt.ssim_window_size, opt.best_auc = ssim_window_size, best_auc
    return opt


def parse_args(argv: Optional[List[str]] = None) -> argparse.Namespace:
    """Flags of src/main.py:207-241 (kept so existing launch scripts parse)."""
    pre = argparse.ArgumentParser(add_help=False)
    pre.add_argument('--config', type=str, default=None)
    pre_args, _ = pre.parse_known_args(argv)
    p = argparse.ArgumentParser(description='Training/Evaluation entrypoint', parents=[pre])
    p.add_argument('--model-type', type=str, default='drct', choices=['drct', 'drn-l'])
    p.add_argument('--dataset', type=str, default='mvtec', choices=['mvtec'])
    p.add_argument('--classe', type=str, default='grid', choices=['grid', 'carpet'])
    p.add_argument('--scale', type=int, default=4, choices=[4, 8])
    p.add_argument('--resolution', type=int, default=128, choices=[32, 64, 128, 256])
    p.add_argument('--epochs', type=int, default=2)
    p.add_argument('--batch-size', type=int, default=4)
    p.add_argument('--lr', type=float, default=2e-4)
    p.add_argument('--no-augment', action='store_true')
    p.add_argument('--device', type=str, default='auto', choices=['auto', 'cuda', 'mps', 'cpu'])
    p.add_argument('--data-root', type=str, default='auto')
    p.add_argument('--save-dir', type=str, default='./workspace/experiment')
    p.add_argument('--pretrain', action='store_true')
    p.add_argument('--test-only', action='store_true')
    p.add_argument('--workers', type=int, default=0 if sys.platform == 'darwin' else 4)
    if pre_args.config is not None and os.path.isfile(pre_args.config):
        import yaml

        with open(pre_args.config, 'r') as f:
            cfg = yaml.safe_load(f) or {}
        p.set_defaults(**{k.replace('-', '_'): v for k, v in cfg.items()})
    return p.parse_args(argv)


def train_drn(opt_drn: DRN) -> None:
    raise NotImplementedError("training is outside the B200 inference+scoring hot path (SURVEY.md 8f row 1)")


def train_drct(opt_drct: DRCT) -> None:
    raise NotImplementedError("training is outside the B200 inference+scoring hot path (SURVEY.md 8f row 1)")
