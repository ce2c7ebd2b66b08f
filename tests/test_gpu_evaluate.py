"""GPU end-to-end parity of the batched evaluator (`python -m src.evaluate` path): PNG folders in the reference's
on-disk layout -> AUC line, against the oracle pipeline (oracle forward + quantise + scoring + AUC) on the same files
and checkpoint.  North-star bar: identical AUC to 1e-3 and the same best SSIM window (or an AUC tie)."""
import importlib
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import drct_oracle as O
from oracle import scoring_oracle as S
from gpu_common import PKG

pytestmark = pytest.mark.gpu


def _write_dataset(root, classe, nc, n):
    hr, lr, labels = S.synthetic_dataset(n, hr=128, nc=nc, scale=4, seed=77)
    for i in range(n):
        split = "good" if labels[i] == 0 else "bad"
        for sub, arr in (("HR", hr[i]), ("LR_4", lr[i])):
            d = os.path.join(root, classe, "test", split, sub)
            os.makedirs(d, exist_ok=True)
            Image.fromarray(arr if nc == 3 else arr[:, :, 0]).save(os.path.join(d, f"{i:03d}.png"))
    return hr, lr, labels


@pytest.mark.parametrize("classe,nc", [("carpet", 3), ("grid", 1)])
def test_evaluate_cli_matches_oracle(tmp_path, classe, nc, capsys):
    evaluate = importlib.import_module(PKG + ".evaluate")
    n = 24
    root = str(tmp_path / "mvtec_128")
    hr, lr, labels = _write_dataset(root, classe, nc, n)
    cfg = O.DrctCfg(n_colors=nc)
    sd = O.make_state_dict(cfg, seed=4)
    ckpt = str(tmp_path / "model_best.pt")
    torch.save(sd, ckpt)

    res = evaluate.main(["--model-type", "drct", "--classe", classe, "--scale", "4", "--resolution", "128", "--data-root",
                         root, "--checkpoint", ckpt, "--batch-size", "16", "--output-dir", str(tmp_path / "out")])
    line = capsys.readouterr().out.strip().splitlines()[-1]
    assert line.startswith("Test AUCs - SSIM(best ws=")

    # oracle pipeline in the evaluator's order: good first (sorted names), then bad
    order = np.argsort(labels, kind="stable")
    x = torch.from_numpy(np.ascontiguousarray(lr[order].transpose(0, 3, 1, 2))).float()
    with torch.no_grad():
        sr = torch.cat([O.drct_forward(sd, x[i:i + 8], cfg) for i in range(0, n, 8)])
    sr_u8 = S.quantize_u8(sr.numpy(), 255.0)
    wss = S.window_sizes_for(128)
    ssim, mse, psnr = S.score_images(list(sr_u8), list(hr[order]), wss)
    y = labels[order]
    best_ws, a_ssim, a_mse, a_psnr = S.aucs_from_scores(y, ssim, mse, psnr, wss)

    assert abs(res["auc_mse"] - a_mse) <= 1e-3 and abs(res["auc_psnr"] - a_psnr) <= 1e-3
    assert abs(res["auc_ssim"] - a_ssim) <= 1e-3
    if res["best_ws"] != best_ws:       # allowed only as an AUC tie between window sizes
        j, jo = wss.index(res["best_ws"]), wss.index(best_ws)
        assert abs(S.roc_auc(y, 1 - ssim[:, j]) - S.roc_auc(y, 1 - ssim[:, jo])) <= 1e-3
    # per-image scores: bf16 noise flips a few uint8 levels, so compare with a tolerance
    assert np.abs(res["scores"][:, :len(wss)] - ssim).max() < 5e-3
    assert np.abs(res["scores"][:, len(wss)] - mse).max() < 1e-3 * max(mse.max(), 1e-3) + 1e-5
    # SR PNGs were written like the reference does (output_dir/<split>/x4/<name>.png)
    assert os.path.isfile(os.path.join(str(tmp_path / "out"), "good", "x4", "000.png"))
