#!/usr/bin/env python
"""Times the 16 x 16-window attention kernel (csrc/attention_tc16.cu) on the five DRCT-L block shapes at the BASELINE
configs[3] size (batch 128, 64 px LR: 2048 windows), next to the mma.sync kernel it replaces.
    python tools/attn16_bench.py [batch]"""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "anomaly-detection-super-resolution_b200"
ops, pack, abi = (importlib.import_module(f"{PKG}.{m}") for m in ("ops", "pack", "_abi"))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    H = W = 64
    ws = 16
    M = B * H * W
    setter = abi.lib().adsr_debug_set_attention_tc
    setter.restype, setter.argtypes = None, [ctypes.c_int]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for C, heads, shift in ((180, 6, 0), (212, 4, 8), (244, 2, 0), (276, 6, 8), (308, 4, 0)):
        hd = C // heads
        hdp = pack.head_pad(hd)
        torch.manual_seed(C)
        qkv = torch.randn(M, 3 * heads * hdp, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, heads * hdp, device="cuda", dtype=torch.bfloat16)
        table = torch.randn((2 * ws - 1) ** 2, heads, device="cuda") * 0.5
        flops = 4.0 * M * ws * ws * heads * hd
        res = {}
        for mode in (1, 0):
            setter(mode)
            ts = []
            for it in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.window_attention(qkv, out, table, B, H, W, ws, shift, heads, hd, hdp)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[mode] = min(ts[1:])
        setter(1)
        print(f"ws16 attention C={C} heads={heads} hd={hd:3d} shift={shift}: tcgen05 {res[1] * 1e3:8.1f} us  {flops / res[1] / 1e9:7.1f} TFLOP/s"
              f"   mma.sync {res[0] * 1e3:8.1f} us  {flops / res[0] / 1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
