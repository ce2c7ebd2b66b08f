"""Debug aid for csrc/conv_halo.cu: identity single-tap weights, input channel 0 = pixel index -> prints which input pixel every
output pixel received."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("anomaly-detection-super-resolution_b200")
ops, pack, abi = (importlib.import_module(f"anomaly-detection-super-resolution_b200.{m}") for m in ("ops", "pack", "_abi"))
H = W = int(os.environ.get("HW", 16)); C = 64; B = 1
x = torch.zeros(B * H * W, C, device="cuda", dtype=torch.bfloat16)
x[:, 0] = torch.arange(H * W, device="cuda").to(torch.bfloat16)
x[:, 1] = 1.0
for tap in [4, 0, 5, 7]:
    w = torch.zeros(C, C, 3, 3, device="cuda")
    w[torch.arange(C), torch.arange(C), tap // 3, tap % 3] = 1.0
    pw = pack.pack_conv3x3_weight(w, None).compact
    out = torch.full((B * H * W, C), -1.0, device="cuda", dtype=torch.bfloat16)
    st = abi.lib().adsr_conv3x3_halo_bf16(abi.ptr(x), C, B, H, W, C, abi.ptr(pw.data), abi.ptr(pw.bias), C, C, 0, 0.0, abi.ptr(out), C, 0, C,
                                          0, abi.num_sms(), abi.stream_ptr())
    torch.cuda.synchronize()
    dy, dx = tap // 3 - 1, tap % 3 - 1
    got = out[:, 0].float().view(H, W)
    ones = out[:, 1].float().view(H, W)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    inside = ((ys + dy >= 0) & (ys + dy < H) & (xs + dx >= 0) & (xs + dx < W)).cuda()
    want = torch.where(inside, ((ys + dy) * W + xs + dx).cuda().float(), torch.zeros((), device="cuda"))
    print(f"tap {tap} (dy {dy} dx {dx}) status {st}: mismatches {int((got != want).sum())} / {H * W}; ch1 ones {int((ones == 1).sum())}")
    if (got != want).any():
        torch.set_printoptions(linewidth=250, precision=0, sci_mode=False)
        print("got[:6]\n", got[:6].int()); print("want[:6]\n", want[:6].int())
