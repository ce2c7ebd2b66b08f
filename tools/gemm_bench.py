"""Per-shape timing of the tcgen05 GEMM / conv kernel and the other DRCT kernels (CUDA events, L2 flushed)."""
import importlib, sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "anomaly-detection-super-resolution_b200"
ops = importlib.import_module(PKG + ".ops"); pack = importlib.import_module(PKG + ".pack")
dev = "cuda"
M = int(os.environ.get("M", 262144))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)

def gemm(K, N, act=0, res=False, name=""):
    a = torch.randn(M, ((K + 63) // 64) * 64, device=dev).to(torch.bfloat16)
    w = pack.pack_gemm_weight(torch.randn(N, K, device=dev) * 0.05, torch.randn(N, device=dev), rows_kernel=bool(int(os.environ.get('ROWS', '1'))))
    out = torch.empty(M, ((N + 63) // 64) * 64, device=dev, dtype=torch.bfloat16)
    r = torch.randn(M, 320, device=dev).to(torch.bfloat16) if res else None
    best, avg = timeit(lambda: ops.tc_gemm(a, K, w, out, act=act, res=r))
    fl = 2.0 * M * K * N
    byt = 2.0 * M * (K + N + (N if res else 0))
    print(f"gemm {name:10s} K={K:4d} N={N:4d} BN={w.BN:3d}x{w.n_tiles} : {best*1e3:8.1f} us  {fl/best/1e9:7.1f} TFLOP/s  {byt/best/1e6:7.1f} GB/s(min traffic)")

_shapes = [(180, 576, 0, False, "qkv1"), (192, 180, 0, True, "proj1"), (180, 360, 2, False, "fc1_1"), (360, 180, 0, True, "fc2_1"), (180, 32, 1, False, "adj1"),
                             (244, 768, 0, False, "qkv3"), (256, 244, 0, True, "proj3"), (244, 488, 2, False, "fc1_3"), (488, 244, 0, True, "fc2_3"),
                             (308, 960, 0, False, "qkv5"), (308, 180, 0, True, "adj5")]
if os.environ.get("GEMM_SHAPE"):
    _shapes = [s_ for s_ in _shapes if s_[4] == os.environ["GEMM_SHAPE"]]
for (K, N, act, res, nm) in _shapes:
    gemm(K, N, act, res, nm)

if os.environ.get("GEMM_ONLY"):
    sys.exit(0)
# attention and LN
B = M // 1024
for (heads, hd, shift) in [(6, 30, 0), (4, 53, 4), (2, 122, 0), (6, 46, 4), (4, 77, 0)]:
    hdp = pack.head_pad(hd)
    qkv = torch.randn(M, 3 * heads * hdp, device=dev).to(torch.bfloat16)
    out = torch.empty(M, heads * hdp, device=dev, dtype=torch.bfloat16)
    table = torch.randn(225, heads, device=dev)
    best, avg = timeit(lambda: ops.window_attention(qkv, out, table, B, 32, 32, 8, shift, heads, hd, hdp))
    fl = 4.0 * M * 64 * heads * hd
    byt = 2.0 * M * 4 * heads * hdp
    print(f"attn heads={heads} hd={hd:3d} shift={shift}: {best*1e3:8.1f} us  {fl/best/1e9:7.1f} TFLOP/s  {byt/best/1e6:7.1f} GB/s")
for C in (180, 244, 308):
    x = torch.randn(M, 320, device=dev).to(torch.bfloat16); o = torch.empty_like(x)
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    best, avg = timeit(lambda: ops.layernorm_rows(x, o, g, b, C))
    print(f"layernorm C={C}: {best*1e3:8.1f} us  {4.0*M*C/best/1e6:7.1f} GB/s")
