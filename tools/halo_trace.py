"""Per-tile timeline of the halo conv's CTA 0 (loader / MMA / epilogue).  Needs a build with HALO_TRACE 1 in csrc/conv_halo.cu."""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
importlib.import_module("anomaly-detection-super-resolution_b200")
ops, pack, abi = (importlib.import_module(f"anomaly-detection-super-resolution_b200.{m}") for m in ("ops", "pack", "_abi"))
B, C, hw = 64, 80, 64
w = torch.randn(C, C, 3, 3, device="cuda") * 0.05
pw = pack.pack_conv3x3_weight(w, torch.randn(C, device="cuda"))
LD = int(os.environ.get("LD", C))
x = torch.randn(B * hw * hw, LD, device="cuda").to(torch.bfloat16)[:, :C]
out = torch.empty(B * hw * hw, LD, device="cuda", dtype=torch.bfloat16)[:, :C]
for _ in range(3):
    ops.conv3x3(x, B, hw, hw, C, pw, out, act=ops.ACT_RELU)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 64 * 2))()
h = ctypes.CDLL(abi.lib()._name)
h.adsr_debug_halo_trace(buf)
t = torch.tensor(list(buf)).view(3, 64, 2)
t0 = int(t[0, 0, 0])
print(" it | a_empty ok  a_full ok | mma wait-start  issue-done | acc_full seen  epi done")
for it in range(15):
    r = [int(t[a, it, k]) - t0 for a in range(3) for k in range(2)]
    print(f"{it:3d} | {r[0]:8d} {r[1]:8d} | {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d}")
