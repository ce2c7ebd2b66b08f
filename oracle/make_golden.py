"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The reference cannot travel to the GPU box, so its outputs on seeded inputs are committed as small
fixtures; the weights are NOT committed, they are regenerated from the seed by
`oracle.drct_oracle.make_state_dict` (a checksum in the fixture guards against RNG drift).

Fixtures written:
  index_maps.npz     window gather maps / attn_mask / relative_position_index from the reference's own
                     torch.roll + window_partition + calculate_mask code (bit-exact objects)
  drct_small.npz     DRCT "small" (embed 60, 4 RDGs) gray x4 16->64 px, B=2: reference SR + taps
  drct_l_rgb.npz     DRCT-L RGB x4 32->128 px, B=1: reference SR output
  scoring.npz        ssim_numpy / psnr_numpy / MSE of the reference on seeded uint8 pairs
  auc.npz            sklearn roc_auc_score on seeded score vectors (ties included)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.drct_oracle import DrctCfg, make_state_dict, state_dict_checksum  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_opt(ref_main, cfg: DrctCfg):
    opt = ref_main.DRCT()
    opt.img_size, opt.n_colors, opt.embed_dim = cfg.img_size, cfg.n_colors, cfg.embed_dim
    opt.depths = (6,) * cfg.num_layers
    opt.num_heads = (cfg.num_heads,) * cfg.num_layers
    opt.window_size, opt.mlp_ratio, opt.upscale = cfg.window_size, cfg.mlp_ratio, cfg.upscale
    opt.scale = [cfg.upscale]
    return opt


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ref_shim.import_reference()
    rdrct, rmetrics, rmain = sys.modules["src.drct"], sys.modules["src.metrics"], sys.modules["src.main"]

    # ---- index maps, from the reference's own view/permute/roll code ------------------------
    maps = {}
    for (H, ws) in [(16, 4), (32, 8), (64, 16), (24, 4)]:
        for shift in (0, ws // 2):
            idx = torch.arange(H * H, dtype=torch.float32).view(1, H, H, 1)
            sh = torch.roll(idx, shifts=(-shift, -shift), dims=(1, 2)) if shift else idx
            win = rdrct.window_partition(sh, ws).view(-1, ws * ws).long()
            maps[f"src_H{H}_ws{ws}_s{shift}"] = win.numpy()
            # inverse path: reverse + roll back must restore the identity
            back = rdrct.window_reverse(win.view(-1, ws, ws, 1).float(), ws, H, H)
            back = torch.roll(back, shifts=(shift, shift), dims=(1, 2)) if shift else back
            assert torch.equal(back.view(-1).long(), torch.arange(H * H))
        blk = rdrct.SwinTransformerBlock(dim=12, input_resolution=(H, H), num_heads=2, window_size=ws,
                                         shift_size=ws // 2)
        maps[f"mask_H{H}_ws{ws}"] = blk.attn_mask.numpy()
        maps[f"rpi_ws{ws}"] = blk.attn.relative_position_index.numpy()
    np.savez_compressed(os.path.join(GOLD, "index_maps.npz"), **maps)

    # ---- DRCT small (C1) with taps ------------------------------------------------------------
    def run_ref(cfg: DrctCfg, x: torch.Tensor, seed: int, jitter: float, hooks: bool):
        sd = make_state_dict(cfg, seed=seed, affine_jitter=jitter)
        model = rdrct.DRCT(ref_opt(rmain, cfg)).eval()
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        assert set(model.state_dict().keys()) == set(sd.keys())
        for k, v in model.state_dict().items():
            assert v.shape == sd[k].shape, k
        taps = {}
        if hooks:
            model.layers[0].swin1.register_forward_hook(lambda m, i, o: taps.__setitem__("l0.swin1", o.detach().numpy()))
            model.layers[0].swin2.register_forward_hook(lambda m, i, o: taps.__setitem__("l0.swin2", o.detach().numpy()))
            model.layers[0].register_forward_hook(lambda m, i, o: taps.__setitem__("l0.out", o.detach().numpy()))
            model.patch_embed.register_forward_hook(lambda m, i, o: taps.__setitem__("embed", o.detach().numpy()))
        with torch.no_grad():
            y = model(x)
        return sd, y, taps

    small = DrctCfg(img_size=16, n_colors=1, embed_dim=60, num_layers=4, num_heads=6, window_size=4)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 1, 16, 16, generator=g) * 255.0
    sd, y, taps = run_ref(small, x, seed=3, jitter=0.1, hooks=True)
    np.savez_compressed(os.path.join(GOLD, "drct_small.npz"), x=x.numpy(), sr=y.numpy(),
                        checksum=state_dict_checksum(sd), **{f"tap.{k}": v for k, v in taps.items()})
    print("drct_small: sr range", float(y.min()), float(y.max()))

    # ---- DRCT-L RGB 32 -> 128, B=1 -----------------------------------------------------------------
    L = DrctCfg(img_size=32, n_colors=3, embed_dim=180, num_layers=12, num_heads=6, window_size=8)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(1, 3, 32, 32, generator=g) * 255.0
    sd, y, _ = run_ref(L, x, seed=1, jitter=0.0, hooks=False)
    np.savez_compressed(os.path.join(GOLD, "drct_l_rgb.npz"), x=x.numpy(), sr=y.numpy().astype(np.float32),
                        checksum=state_dict_checksum(sd))
    print("drct_l_rgb: sr range", float(y.min()), float(y.max()), "params", sum(v.numel() for v in sd.values()))

    # ---- scoring: reference ssim_numpy / psnr_numpy on seeded uint8 pairs ----------------------------
    rng = np.random.default_rng(5)
    cases = {}
    k = 0
    for (H, W, C) in [(24, 24, 1), (24, 32, 3), (40, 40, 3)]:
        for variant in range(3):
            hr = rng.integers(0, 256, size=(H, W, C), dtype=np.uint8)
            if variant == 0:
                sr = np.clip(hr.astype(np.int32) + rng.integers(-12, 13, size=hr.shape), 0, 255).astype(np.uint8)
            elif variant == 1:
                sr = hr.copy()                        # mse == 0 -> psnr inf path (metrics.py:21-22)
            else:
                sr = np.full_like(hr, 128)            # constant image
            wss = [w for w in range(3, max(3, min(H, W) - 3) + 1, 10) if w % 2 == 1] or [3]
            hr_f, sr_f = hr.astype(np.float32) / 255.0, sr.astype(np.float32) / 255.0
            ss = [rmetrics.ssim_numpy(hr_f, sr_f, w) for w in wss]
            d = sr_f - hr_f
            cases[f"c{k}.hr"], cases[f"c{k}.sr"] = hr, sr
            cases[f"c{k}.ws"], cases[f"c{k}.ssim"] = np.asarray(wss), np.asarray(ss, dtype=np.float64)
            cases[f"c{k}.mse"] = np.float64(np.mean(d * d))
            cases[f"c{k}.psnr"] = np.float64(rmetrics.psnr_numpy(hr_f, sr_f))
            k += 1
    cases["n"] = np.int64(k)
    np.savez_compressed(os.path.join(GOLD, "scoring.npz"), **cases)
    print("scoring cases", k)

    # ---- DRN: reference src/drn.py on oracle-generated weights (key/shape layout checked strictly) ---------------------
    from oracle import drn_oracle as DO
    rdrn = sys.modules["src.drn"]
    for name, cfg, B in (("drn_l_rgb", DO.DrnCfg(), 1), ("drn_small_gray", DO.DrnCfg(n_blocks=3, n_colors=1), 2)):
        opt = rmain.DRN()
        opt.scale = [2 ** (i + 1) for i in range(cfg.phase)]
        opt.n_blocks, opt.n_feats, opt.n_colors, opt.rgb_range, opt.negval = cfg.n_blocks, cfg.n_feats, cfg.n_colors, 255, cfg.negval
        model = rdrn.DRN(opt).eval()
        sd = DO.make_state_dict(cfg, seed=5)
        assert set(model.state_dict().keys()) == set(sd.keys()), set(model.state_dict().keys()) ^ set(sd.keys())
        for kk, v in model.state_dict().items():
            assert v.shape == sd[kk].shape, kk
        model.load_state_dict(sd, strict=True)
        gx = torch.Generator().manual_seed(13)
        x = torch.rand(B, cfg.n_colors, 32, 32, generator=gx) * 255.0
        with torch.no_grad():
            ys = model(x)
        np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), x=x.numpy(), checksum=DO.state_dict_checksum(sd),
                            **{f"sr{i}": y.numpy() for i, y in enumerate(ys)})
        print(name, [tuple(y.shape) for y in ys], float(ys[-1].min()), float(ys[-1].max()))

    # ---- psnr_torch / ssim_torch (validation metrics, src/metrics.py:70-108) and float-input ssim/psnr_numpy --------
    gt = torch.Generator().manual_seed(21)
    mt = {}
    k = 0
    for (C, H, W, rr) in [(3, 40, 40, 255.0), (1, 24, 24, 255.0), (3, 32, 48, 1.0), (1, 8, 8, 255.0)]:
        hr = torch.rand(1, C, H, W, generator=gt) * rr
        sr = (hr + (torch.rand(1, C, H, W, generator=gt) - 0.5) * 0.2 * rr)
        mt[f"t{k}.hr"], mt[f"t{k}.sr"], mt[f"t{k}.rgb_range"] = hr.numpy(), sr.numpy(), np.float64(rr)
        mt[f"t{k}.psnr"] = np.float64(rmetrics.psnr_torch(sr, hr, rr))
        mt[f"t{k}.ssim"] = np.float64(rmetrics.ssim_torch(sr, hr, rr))
        mt[f"t{k}.ssim7"] = np.float64(rmetrics.ssim_torch(sr, hr, rr, win_size=7))
        a = (hr[0].permute(1, 2, 0).numpy() / rr).astype(np.float32)
        b = (sr[0].permute(1, 2, 0).numpy() / rr).astype(np.float32)
        mt[f"t{k}.np_ssim5"] = np.float64(rmetrics.ssim_numpy(a, b, 5))
        mt[f"t{k}.np_psnr"] = np.float64(rmetrics.psnr_numpy(a, b))
        mt[f"t{k}.np_ssim5_dr255"] = np.float64(rmetrics.ssim_numpy(a * 255, b * 255, 5, data_range=255.0))
        k += 1
    mt["n"] = np.int64(k)
    np.savez_compressed(os.path.join(GOLD, "metrics_api.npz"), **mt)

    # ---- AUC -------------------------------------------------------------------------------------------
    from sklearn.metrics import roc_auc_score

    rng = np.random.default_rng(9)
    auc = {}
    for i in range(6):
        n = 40 + 13 * i
        y_true = (rng.random(n) < 0.5).astype(np.int64)
        y_true[0], y_true[1] = 0, 1
        s = rng.normal(size=n) + 0.7 * y_true
        if i % 2:
            s = np.round(s, 1)                      # force ties
        auc[f"y{i}"], auc[f"s{i}"], auc[f"auc{i}"] = y_true, s, np.float64(roc_auc_score(y_true, s))
    auc["n"] = np.int64(6)
    np.savez_compressed(os.path.join(GOLD, "auc.npz"), **auc)
    print("done ->", GOLD)


if __name__ == "__main__":
    main()
