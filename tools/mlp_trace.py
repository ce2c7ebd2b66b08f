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
stats = torch.zeros(M, 2, 2, device=dev)
yf = y[:, :C].float(); stats[:, 0, 0] = yf.sum(1); stats[:, 0, 1] = (yf * yf).sum(1)
ops.swin_mlp(y, C, pm, z, stats_in=(stats, 2)); torch.cuda.synchronize()
trace = torch.zeros(3, 8, 64, 8, dtype=torch.int64, device=dev)
fn = abi.lib().adsr_debug_set_mlp_trace; fn.restype = None; fn.argtypes = [ctypes.c_void_p]
fn(trace.data_ptr())
ops.swin_mlp(y, C, pm, z, stats_in=(stats, 2)); torch.cuda.synchronize()
fn(None)
t = trace.cpu()
t0 = int(t[t > 0].min())
plan = pm.plan.tolist(); ns = plan[16]; npro = plan[17]
st = [plan[18 + 8 * i: 26 + 8 * i] for i in range(ns)]
rel = lambda v: int(v) - t0 if int(v) > 0 else -1
for it in range(4):
    print(f"---- tile {it}")
    for i in range(ns):
        b, rows, ks, kind, ch, kidx, dcol, fl = st[i]
        print(f" st{i:2d} {'fc1' if kind == 0 else 'fc2'} ch{ch} k{kidx} N={rows:3d} fl={fl:2d} | loader wait {rel(t[0,it,i,0]):7d} ->{rel(t[0,it,i,1]):7d}"
              f" | mma start {rel(t[1,it,i,0]):7d} h_ok {rel(t[1,it,i,1]):7d} w_ok {rel(t[1,it,i,2]):7d} issued {rel(t[1,it,i,3]):7d}")
    for j in range(plan[1]):
        print(f" epi1 ch{j}: wait {rel(t[2,it,j,0]):7d} acc_ok {rel(t[2,it,j,1]):7d} math_done {rel(t[2,it,j,2]):7d} bar {rel(t[2,it,j,3]):7d} done {rel(t[2,it,j,4]):7d}")
    print(f" epi2: wait {rel(t[2,it,16,0]):7d} acc2_ok {rel(t[2,it,16,1]):7d} done {rel(t[2,it,16,2]):7d}")
