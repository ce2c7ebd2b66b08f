"""Times DRN-L's RCAB conv (80 -> 80, batch 64) at 32^2 and 64^2 through the halo-tile kernel and the streaming implicit GEMM."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
importlib.import_module("anomaly-detection-super-resolution_b200")
ops, pack = (importlib.import_module(f"anomaly-detection-super-resolution_b200.{m}") for m in ("ops", "pack"))
B, C = 64, int(os.environ.get("C", 80))
iters = int(os.environ.get("ITERS", 20))
torch.manual_seed(0)
w = torch.randn(C, C, 3, 3, device="cuda") * 0.05
pw = pack.pack_conv3x3_weight(w, torch.randn(C, device="cuda"))
for hw in (32, 64):
    m = B * hw * hw
    LD = int(os.environ.get("LD", C))
    x = torch.randn(m, LD, device="cuda").to(torch.bfloat16)[:, :C]
    out = torch.empty(m, LD, device="cuda", dtype=torch.bfloat16)[:, :C]
    for halo in (True, False):
        ops._HALO_CONV = halo
        for _ in range(3):
            ops.conv3x3(x, B, hw, hw, C, pw, out, act=ops.ACT_RELU)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv3x3(x, B, hw, hw, C, pw, out, act=ops.ACT_RELU)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        print(f"{hw}x{hw} halo={int(halo)}: {us:7.1f} us  {2.0 * m * 9 * C * C / us / 1e6:7.1f} TFLOP/s")
