"""GPU parity of the first training-step slice against torch autograd / torch.optim.Adam in fp32:
L1 loss + gradient (src/loss.py:83-84, 108-121), conv_last backward (src/drct.py:847, 895), fused Adam (src/trainer.py:49-59)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_common import mod, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("shape", [(4, 3, 128, 128), (2, 1, 33, 17), (1, 3, 5, 7)])
def test_l1_loss_and_grad(shape):
    tr = mod("training")
    torch.manual_seed(sum(shape))
    sr = (torch.rand(shape, device=DEV) * 255).requires_grad_(True)
    hr = torch.rand(shape, device=DEV) * 255
    with torch.no_grad():
        sr.view(-1)[::7] = hr.view(-1)[::7]                 # exact ties: sign(0) = 0 like torch
    want = F.l1_loss(sr, hr)
    want.backward()
    loss, grad = tr.l1_loss_and_grad(sr.detach(), hr, grad_scale=1.0)
    assert abs(float(loss) - float(want.detach())) <= 1e-5 * max(1.0, abs(float(want.detach())))
    assert torch.equal(grad, sr.grad)
    _, grad2 = tr.l1_loss_and_grad(sr.detach(), hr, grad_scale=128.0)     # e.g. a loss scale
    assert torch.allclose(grad2, sr.grad * 128.0, rtol=1e-6, atol=0)


@pytest.mark.parametrize("B,H,W,nc", [(2, 128, 128, 3), (3, 24, 40, 1), (1, 8, 8, 3)])
def test_conv_last_backward_vs_autograd(B, H, W, nc):
    tr = mod("training")
    torch.manual_seed(B + H + nc)
    x = (torch.randn(B * H * W, 64, device=DEV) * 0.7).to(torch.bfloat16)
    w = (torch.randn(nc, 64, 3, 3, device=DEV) * 0.1).requires_grad_(True)
    bias = torch.zeros(nc, device=DEV, requires_grad=True)
    g = torch.randn(B, nc, H, W, device=DEV) * 0.01
    xi = x.float().view(B, H, W, 64).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    y = F.conv2d(xi, w, bias, padding=1)
    y.backward(g)
    dx, dw, db = tr.conv_last_backward(x, B, H, W, w.detach(), g)
    dx_want = xi.grad.permute(0, 2, 3, 1).reshape(B * H * W, 64)
    assert rel_err(dx, dx_want) < 6e-3                       # bf16 storage of dx
    assert rel_err(dw, w.grad) < 2e-4 and rel_err(db, bias.grad) < 2e-4
    dx2, dw2, db2 = tr.conv_last_backward(x, B, H, W, w.detach(), g)          # fixed reduction order: bit-identical
    assert torch.equal(dw, dw2) and torch.equal(db, db2) and torch.equal(dx, dx2)
    none_dx, _, _ = tr.conv_last_backward(x, B, H, W, w.detach(), g, need_dx=False)
    assert none_dx is None


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_matches_torch(wd):
    tr = mod("training")
    torch.manual_seed(3)
    shapes = [(180, 180), (540,), (3, 64, 3, 3), (70000,), (1,), (961, 6)]
    ps = [torch.randn(s, device=DEV) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt_ref = torch.optim.Adam(ref, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    opt = tr.FusedAdam(ps, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    for step in range(6):
        grads = [torch.randn_like(p) * (0.1 + step) for p in ps]
        for r, g in zip(ref, grads):
            r.grad = g.clone()
        opt_ref.step()
        opt.step(grads)
    for p, r in zip(ps, ref):
        assert torch.allclose(p, r.detach(), rtol=2e-6, atol=2e-7), float((p - r.detach()).abs().max())


def test_one_training_step_of_the_last_layer():
    """forward of conv_last -> L1 loss + gradient -> conv_last backward -> fused Adam on its parameters: one optimisation step
    equals torch's (conv2d + l1_loss + autograd + torch.optim.Adam, fp32).  The first Adam step moves every weight by
    lr * sign(gradient), so a weight whose gradient is ~0 may legitimately land 2 * lr apart: those are counted, not compared."""
    tr, ops = mod("training"), mod("ops")
    torch.manual_seed(11)
    B, H, W, nc, lr = 2, 64, 64, 3, 1e-3
    x = (torch.randn(B * H * W, 64, device=DEV) * 0.5).to(torch.bfloat16)
    w, bias = torch.randn(nc, 64, 3, 3, device=DEV) * 0.05, torch.randn(nc, device=DEV) * 0.1
    hr = torch.rand(B, nc, H, W, device=DEV)
    sr = torch.empty(B, nc, H, W, device=DEV)
    ops.conv_last_quant(x, B, H, W, 64, w.contiguous(), bias, nc, torch.zeros(nc, device=DEV), 1.0, 1.0, sr, None)
    loss, g = tr.l1_loss_and_grad(sr, hr)
    _, dw, db = tr.conv_last_backward(x, B, H, W, w, g, need_dx=False)
    w_new, b_new = w.clone(), bias.clone()
    tr.FusedAdam([w_new, b_new], lr=lr).step([dw, db])

    wr, br = w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    xi = x.float().view(B, H, W, 64).permute(0, 3, 1, 2)
    y = F.conv2d(xi, wr, br, padding=1)
    assert float((sr - y.detach()).abs().max()) < 1e-4
    loss_ref = F.l1_loss(y, hr)
    loss_ref.backward()
    torch.optim.Adam([wr, br], lr=lr).step()
    assert abs(float(loss) - float(loss_ref)) < 1e-5
    assert rel_err(dw, wr.grad) < 1e-3 and rel_err(db, br.grad) < 1e-3
    far = ((w_new - wr.detach()).abs() > 1e-5).float().mean()
    assert float(far) < 0.01, f"{float(far):.4f} of the weights moved differently"
    assert float((b_new - br.detach()).abs().max()) < 1e-5
