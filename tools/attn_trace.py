"""Timeline of CTA 0 of the fused attention-half kernel (clock64 at the hand-over points), first 4 tiles.
usage: C=212 HEADS=4 SHIFT=4 python tools/attn_trace.py"""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "anomaly-detection-super-resolution_b200"
ops = importlib.import_module(PKG + ".ops"); pack = importlib.import_module(PKG + ".pack"); abi = importlib.import_module(PKG + "._abi")
dev = "cuda"
C, heads, shift = int(os.environ.get("C", 212)), int(os.environ.get("HEADS", 4)), int(os.environ.get("SHIFT", 4))
B, H, W = int(os.environ.get("B", 256)), 32, 32
M = B * H * W
hd = C // heads; hdp = pack.head_pad(hd)
x = torch.randn(M, 320, device=dev).to(torch.bfloat16); y = torch.empty_like(x)
pa = pack.pack_swin_attn(torch.randn(3 * C, C, device=dev) * 0.05, torch.randn(3 * C, device=dev) * 0.1, torch.ones(C, device=dev),
                         torch.zeros(C, device=dev), 1e-5, torch.randn(C, C, device=dev) * 0.05, torch.randn(C, device=dev) * 0.1, heads)
table = torch.randn(225, heads, device=dev) * 0.2
stats = torch.zeros(M, 2, 2, device=dev); xf = x[:, :C].float(); stats[:, 0, 0] = xf.sum(1); stats[:, 0, 1] = (xf * xf).sum(1)
st_y = torch.zeros(M, 8, 2, device=dev)
att = torch.empty(M, heads * hdp, device=dev, dtype=torch.bfloat16)
mode = ops.swin_attn_mode(C, heads, hdp, True)
run = (lambda: ops.swin_attn(x, pa, table, y, B, H, W, shift, (stats, 2), True, stats_out=(st_y, 0))) if mode == 2 else \
      (lambda: ops.swin_attn(x, pa, table, att, B, H, W, shift, (stats, 2), False))
run(); torch.cuda.synchronize()
trace = torch.zeros(3 * 4 * 9 * 8, dtype=torch.int64, device=dev)
setter = abi.lib().adsr_debug_set_attn_trace
setter.restype, setter.argtypes = None, [ctypes.c_void_p]
setter(trace.data_ptr()); run(); torch.cuda.synchronize(); setter(None)
t = trace.cpu().view(3, 4, 9, 8)
t0 = int(t[t > 0].min())
rel = lambda v: f"{(int(v) - t0):7d}" if int(v) > 0 else "      -"
names = {0: ["top", "qkv(g+1) all issued", "S issue", "PV issue", "PV issued", "", "", ""],
         1: ["wait qkv_full", "qkv_full seen", "qkv epi done", "s_full seen", "softmax done", "o_full seen", "O epi done", ""]}
print(f"C={C} heads={heads} hdp={hdp} shift={shift} mode={mode}   (SM cycles relative to the first event)")
for it in range(3):
    print(f"--- tile {it}")
    print("  x loader: wait x_empty %s  got %s  issued %s" % tuple(rel(v) for v in t[2, it, 8, :3]))
    print("  mma tile end: proj start %s  proj issued %s  next qkv issued %s" % tuple(rel(v) for v in t[0, it, 8, :3]))
    for h in range(heads):
        print(f"  head {h} mma: " + "  ".join(f"{names[0][k]} {rel(t[0, it, h, k])}" for k in range(5)))
        if int(t[2, it, h, 4]) > 0:
            print(f"  head {h} qkv issuer (x-loader warp): wait {rel(t[2, it, h, 3])}  go {rel(t[2, it, h, 4])}  all issued {rel(t[2, it, h, 5])}")
        print(f"  head {h} epi: " + "  ".join(f"{names[1][k]} {rel(t[1, it, h, k])}" for k in range(7)))
    print("  epi: wait proj_full %s  seen %s  proj epi done %s" % tuple(rel(v) for v in t[1, it, 8, :3]))
