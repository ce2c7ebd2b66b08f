"""Dumps the CTA-0 timeline of one fused-MLP launch (debug hook adsr_debug_set_mlp_trace)."""
import importlib, sys, os, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "anomaly-detection-super-resolution_b200"
ops = importlib.import_module(PKG + ".ops"); pack = importlib.import_module(PKG + ".pack"); abi = importlib.import_module(PKG + "._abi")
dev = "cuda"
M = int(os.environ.get("M", 262144)); C = int(os.environ.get("C", 180)); H = int(os.environ.get("H", 360))
y = torch.randn(M, 320, device=dev).to(torch.bfloat16); z = torch.empty_like(y)
pm = pack.pack_swin_mlp(torch.randn(H, C, device=dev) * 0.05, torch.randn(H, device=dev), torch.ones(C, device=dev),
                        torch.zeros(C, device=dev), 1e-5, torch.randn(C, H, device=dev) * 0.05, torch.randn(C, device=dev))
FOLD = os.environ.get("FOLD")            # FOLD=1: the folded-adjust variant (swin_mlp_adjust), unset: plain MLP
if FOLD == "res":                     # FOLD=res: adjust5 + residual folded in (swin_mlp_conv_res), C=308 H=308
    pm = pack.pack_swin_mlp_conv_res(torch.randn(H, C, device=dev) * 0.05, torch.randn(H, device=dev), torch.ones(C, device=dev),
                                     torch.zeros(C, device=dev), 1e-5, torch.randn(C, H, device=dev) * 0.05, torch.randn(C, device=dev),
                                     torch.randn(180, C, device=dev) * 0.05, torch.randn(180, device=dev), 0.2)
    slab = torch.randn(M, 320, device=dev).to(torch.bfloat16); st_s = torch.zeros(M, 12, 2, device=dev)
elif FOLD is not None:
    pm = pack.pack_swin_mlp(torch.randn(H, C, device=dev) * 0.05, torch.randn(H, device=dev), torch.ones(C, device=dev),
                            torch.zeros(C, device=dev), 1e-5, torch.randn(C, H, device=dev) * 0.05, torch.randn(C, device=dev),
                            torch.randn(32, C, device=dev) * 0.05, torch.randn(32, device=dev))
    slab = torch.zeros(M, 320, device=dev, dtype=torch.bfloat16); st_s = torch.zeros(M, 12, 2, device=dev)
stats = torch.zeros(M, 2, 2, device=dev)
yf = y[:, :C].float(); stats[:, 0, 0] = yf.sum(1); stats[:, 0, 1] = (yf * yf).sum(1)
run = (lambda: ops.swin_mlp(y, C, pm, z, stats_in=(stats, 2))) if FOLD is None else \
    (lambda: ops.swin_mlp_conv_res(y, C, pm, slab, slab, stats_in=(stats, 2), stats_out=(st_s, 0))) if FOLD == "res" else \
    (lambda: ops.swin_mlp_adjust(y, C, pm, slab, C, stats_in=(stats, 2), stats_out=(st_s, 2)))
run(); torch.cuda.synchronize()
trace = torch.zeros(4, 8, 64, 8, dtype=torch.int64, device=dev)
fn = abi.lib().adsr_debug_set_mlp_trace; fn.restype = None; fn.argtypes = [ctypes.c_void_p]
fn(trace.data_ptr())
run(); torch.cuda.synchronize()
fn(None)
t = trace.cpu()
t0 = int(t[t > 0].min())
nc = pm.plan.tolist()[2]
rel = lambda v: int(v) - t0 if int(v) > 0 else -1
for it in range(1, 5):
    print(f"---- tile {it}")
    print(f" tile warp: z wait {rel(t[0,it,0,0]):7d} z_ready {rel(t[0,it,0,1]):7d} store read {rel(t[0,it,0,2]):7d} load(it+2) issued {rel(t[0,it,0,3]):7d}"
          f" | fc1 a_full: wait {rel(t[0,it,1,0]):7d} ok {rel(t[0,it,1,1]):7d}")
    for j in range(nc):
        print(f" fc1 ch{j}: start {rel(t[1,it,j,0]):7d} acc1_free {rel(t[1,it,j,1]):7d} issued {rel(t[1,it,j,2]):7d}"
              f" | epi1: wait {rel(t[2,it,j,0]):7d} acc_ok {rel(t[2,it,j,1]):7d} slab0 {rel(t[2,it,j,2]):7d} slab1 {rel(t[2,it,j,3]):7d}")
        for s in range(2):
            print(f"   fc2 ch{j} slab{s}: start {rel(t[3,it,2*j+s,0]):7d} h_ok {rel(t[3,it,2*j+s,1]):7d} w2_full {rel(t[3,it,2*j+s,5]):7d} issued {rel(t[3,it,2*j+s,2]):7d}"
                  + (f"   (h_ready {rel(t[3,it,0,3]):7d} acc2_free {rel(t[3,it,0,4]):7d})" if j == 0 and s == 0 else ""))
    print(f" epi3 (adjust columns): wait {rel(t[2,it,17,0]):7d} acc_ok {rel(t[2,it,17,1]):7d} done {rel(t[2,it,17,2]):7d}")
    print(f" epi2: wait {rel(t[2,it,16,0]):7d} acc2_ok {rel(t[2,it,16,1]):7d} done {rel(t[2,it,16,2]):7d}")
