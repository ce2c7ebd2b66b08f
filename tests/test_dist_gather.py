"""CPU, world_size 2, gloo: the multi-rank part of the evaluator (image index -> rank i mod R sharding, padded
all_gather of score rows + ids, id-ordered table on rank 0, best-ws / AUC selection) -- SURVEY.md section 8e."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "anomaly-detection-super-resolution_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, n_ws, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    evaluate = importlib.import_module(PKG + ".evaluate")
    full = torch.from_numpy(np.random.default_rng(0).random((n_total, n_ws + 2)))
    ids = torch.arange(rank, n_total, world, dtype=torch.int64)          # the evaluator's sharding rule
    table = evaluate.gather_scores(full[ids].clone(), ids, n_total)
    # the sync-free form: every rank gets the gathered device rows, the host table is built when it is consumed
    dev_rows = evaluate.gather_scores(full[ids].clone(), ids, n_total, to_host=False)
    assert isinstance(dev_rows, torch.Tensor) and dev_rows.shape == (world * ((n_total + world - 1) // world), n_ws + 3)
    assert np.array_equal(evaluate.scores_table(dev_rows, n_total), full.numpy())
    if rank == 0:
        np.save(out_path, table)
    else:
        assert table is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 7, 1])     # 7: ranks hold different counts -> padded gather; 1: a rank with no image
def test_gather_scores_two_ranks(tmp_path, n_total):
    n_ws = 3
    out = str(tmp_path / "table.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_total, n_ws, out), nprocs=2, join=True)
    want = np.random.default_rng(0).random((n_total, n_ws + 2))
    assert np.array_equal(np.load(out), want)


def test_aucs_from_scores_matches_oracle():
    from oracle import scoring_oracle as S
    metrics = importlib.import_module(PKG + ".metrics")
    rng = np.random.default_rng(3)
    n, wss = 60, [3, 13, 23]
    y = (np.arange(n) >= n // 2).astype(int)
    ssim = np.clip(0.8 - 0.1 * y[:, None] + 0.1 * rng.normal(size=(n, 3)), 0, 1)
    mse = np.abs(rng.normal(size=n)) * (1 + y)
    psnr = 10 * np.log10(1.0 / np.maximum(mse, 1e-9))
    table = np.concatenate([ssim, mse[:, None], psnr[:, None]], 1)
    got = metrics.aucs_from_scores(y, table, wss)
    want = S.aucs_from_scores(y, ssim, mse, psnr, wss)
    assert got[0] == want[0] and np.allclose(got[1:], want[1:], atol=1e-12)


def _grad_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    training = importlib.import_module(PKG + ".training")
    g = torch.Generator().manual_seed(100 + rank)
    grads = [torch.randn(s, generator=g) for s in ((7, 5), (13,), (2, 3, 3, 3))]
    training.allreduce_gradients(grads)
    if rank == 0:
        torch.save(grads, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_gradients_two_ranks(tmp_path):
    """Data-parallel gradient exchange of the training step (one flat bucket, averaged): world size 2 on gloo."""
    out = str(tmp_path / "grads.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    want = []
    for s in ((7, 5), (13,), (2, 3, 3, 3)):
        want.append(None)
    gens = [torch.Generator().manual_seed(100 + r) for r in range(2)]
    per_rank = [[torch.randn(s, generator=g) for s in ((7, 5), (13,), (2, 3, 3, 3))] for g in gens]
    for i, gt in enumerate(got):
        assert torch.allclose(gt, (per_rank[0][i] + per_rank[1][i]) / 2, atol=1e-7)

