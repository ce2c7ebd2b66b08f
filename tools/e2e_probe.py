"""Times the evaluator's two host-input paths (synchronous step vs. double-buffered submit/step_submitted)."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
PKG = bench.PKG
drct = importlib.import_module(PKG + ".drct"); evaluate = importlib.import_module(PKG + ".evaluate")
metrics = importlib.import_module(PKG + ".metrics"); main_mod = importlib.import_module(PKG + ".main")
dev = torch.device("cuda", 0); B = 256
opt = main_mod.setup_opt_drct(main_mod.DRCT(), 0.0, 11, "mvtec", "carpet", False, 4, True, 3, 1, B, 128, 32, "", "", "", 1, 1, 1, 0.0, 0, ".", "1*L1")
torch.manual_seed(1)
model = drct.DRCT(opt).to(dev).eval()
hr_u8, lr_u8, _ = bench.synthetic_pairs(64)
lr_h = bench.to_float_nchw(np.tile(lr_u8, (4, 1, 1, 1))).pin_memory(); hr_h = bench.to_float_nchw(np.tile(hr_u8, (4, 1, 1, 1))).pin_memory()
ev = evaluate.BatchedEvaluator(model, 255.0, metrics.window_sizes_for(128))
lr_d, hr_d = lr_h.to(dev), hr_h.to(dev)
K = 10
def run(name, fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name:28s} {1e3*dt/K:7.2f} ms/step  {B*K/dt:8.1f} img/s")
def dev_only():
    for _ in range(K): ev.step(lr_d, hr_d).cpu()
def sync():
    for _ in range(K): ev.step(lr_h, hr_h).cpu()
def piped():
    nxt = ev.submit(lr_h, hr_h)
    for _ in range(K):
        cur, nxt = nxt, ev.submit(lr_h, hr_h)
        ev.step_submitted(cur).cpu()
def copy_only():
    for _ in range(K): lr_h.to(dev, non_blocking=True); hr_h.to(dev, non_blocking=True); torch.cuda.synchronize()
for _ in range(2):
    run("device-resident inputs", dev_only); run("host inputs, in-stream copy", sync); run("host inputs, double buffer", piped); run("copies only", copy_only)
