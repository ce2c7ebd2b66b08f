"""Timing of the window attention kernels per DRCT block shape (tcgen05 kernel vs the mma.sync kernel)."""
import importlib, sys, os, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "anomaly-detection-super-resolution_b200"
ops = importlib.import_module(PKG + ".ops"); pack = importlib.import_module(PKG + ".pack"); abi = importlib.import_module(PKG + "._abi")
dev = "cuda"
M = int(os.environ.get("M", 262144))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
setter = abi.lib().adsr_debug_set_attention_tc; setter.restype = None; setter.argtypes = [ctypes.c_int]

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

B = M // 1024
shapes = [(6, 30, 0), (4, 53, 4), (2, 122, 0), (6, 46, 4), (4, 77, 0)]
if os.environ.get('ATTN_ONLY'):
    shapes = [shapes[int(os.environ['ATTN_ONLY'])]]
for (heads, hd, shift) in shapes:
    hdp = pack.head_pad(hd)
    qkv = torch.randn(M, 3 * heads * hdp, device=dev).to(torch.bfloat16)
    out = torch.empty(M, heads * hdp, device=dev, dtype=torch.bfloat16)
    table = torch.randn(225, heads, device=dev)
    res = []
    for tc in (2, 0):
        setter(tc)
        res.append(timeit(lambda: ops.window_attention(qkv, out, table, B, 32, 32, 8, shift, heads, hd, hdp)))
    setter(1)
    byt = 2.0 * M * 4 * heads * hdp
    print(f"attn heads={heads} hd={hd:3d} shift={shift}: tcgen05 {res[0]*1e3:8.1f} us ({byt/res[0]/1e6:7.1f} GB/s)   mma.sync {res[1]*1e3:8.1f} us ({byt/res[1]/1e6:7.1f} GB/s)")
