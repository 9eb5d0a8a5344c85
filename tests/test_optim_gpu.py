"""GPU: KD-loss backward kernel vs autograd of the oracle loss, and the fused flat AdamW (+ loss scaling, global-norm
clipping, skip-on-inf, dynamic scale) vs torch.optim.AdamW + clip_grad_norm_ driven the way the reference's
NativeScalerWithGradNormCount drives them (stage1/utils.py:341-368).  Tolerances: fp32 round-off (different op order)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,E,img,w", [(3, 64, 12, 192, 1.0), (2, 1024, 18, 252, 0.5), (1, 32, 7, 100, 0.0)])
def test_kd_loss_backward_matches_autograd(cuda, B, C, E, img, w):
    from efficientsam3_b200 import ops
    from efficientsam3_b200.stage1.optim import KDLossFunction
    from oracle import kd_loss as O
    g = torch.Generator().manual_seed(B * 100 + E)
    preds = torch.randn(B, C, E, E, generator=g)
    teach = torch.randn(B, C, E, E, generator=g)
    sizes = [(3, img, img * 3 // 4) if i % 2 == 0 else (3, img * 2 // 3, img) for i in range(B)]
    p = preds.clone().requires_grad_(True)
    loss, _, _ = O.kd_loss(p, teach, img, sizes, w)
    (loss * 3.0).backward()
    sz = torch.tensor([[s[1], s[2]] for s in sizes], dtype=torch.int32, device=cuda)
    out, per = ops.kd_loss_fwd(preds.to(cuda), teach.to(cuda), sz, img, w)
    got = ops.kd_loss_bwd(preds.to(cuda), teach.to(cuda), sz, per, img, w, grad_scale=3.0).cpu()
    scale = p.grad.abs().max().item()
    assert (got - p.grad).abs().max().item() <= 2e-5 * scale, (got - p.grad).abs().max().item() / scale
    assert (got == 0).sum().item() == (p.grad == 0).sum().item()          # padded pixels get exactly zero gradient
    # through autograd, with the device-resident loss scale folded in by the same kernel
    pc = preds.to(cuda).requires_grad_(True)
    l2 = KDLossFunction.apply(pc, teach.to(cuda), sz, img, w)
    assert abs(l2.item() - loss.item()) <= 1e-5 * abs(loss.item())
    (l2 * 3.0).backward()
    assert (pc.grad.cpu() - p.grad).abs().max().item() <= 2e-5 * scale


def _models(cuda):
    torch.manual_seed(3)
    mk = lambda: torch.nn.Sequential(torch.nn.Conv2d(3, 6, 3, padding=1), torch.nn.BatchNorm2d(6), torch.nn.Flatten(),
                                     torch.nn.Linear(6 * 25, 33), torch.nn.LayerNorm(33), torch.nn.Linear(33, 5, bias=False))
    a = mk()
    b = mk()
    b.load_state_dict(a.state_dict())
    return a.to(cuda), b.to(cuda)


def test_flat_adamw_matches_torch_adamw_with_scaling_and_clipping(cuda):
    from efficientsam3_b200.stage1.optim import FlatAdamW, split_decay
    ma, mb = _models(cuda)
    S, world = 1024.0, 2
    opt = FlatAdamW(ma, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05, loss_scale=S)
    d, n = split_decay(mb.named_parameters())
    ref = torch.optim.AdamW([{"params": [p for _, p in d]}, {"params": [p for _, p in n], "weight_decay": 0.0}], lr=1e-2,
                            betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    g = torch.Generator().manual_seed(9)
    norms = []
    for step in range(6):
        lr = 1e-2 * (0.7 ** step)
        mag = 10.0 if step % 2 == 0 else 0.01                          # clipping engages on the large steps only
        for (_, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            gr = (torch.randn(pb.shape, generator=g) * mag).to(cuda)
            pb.grad = gr.clone()
            pa.grad.copy_(gr * S * world)                             # loss-scaled and summed over `world` ranks
        for grp in ref.param_groups:
            grp["lr"] = lr
        norms.append(torch.nn.utils.clip_grad_norm_(mb.parameters(), 5.0).item())
        ref.step()
        opt.step(lr=lr, max_norm=5.0, world_size=world)
        assert abs(opt.last_grad_norm() - norms[-1]) <= 1e-4 * norms[-1]
    assert opt.step_count() == 6 and float(opt.state[0]) == S
    for (na, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        err = (pa - pb).abs().max().item()
        assert err <= 2e-5 * max(pb.abs().max().item(), 1e-3), (na, err)
    assert norms[0] > 5.0 > norms[1]


def test_flat_adamw_skips_on_inf_and_adapts_the_scale(cuda):
    from efficientsam3_b200.stage1.optim import FlatAdamW
    ma, _ = _models(cuda)
    opt = FlatAdamW(ma, lr=1e-2, loss_scale=65536.0, dynamic_loss_scale=True, growth_interval=3)
    before = opt.flat_param.clone()
    opt.flat_grad.normal_()
    opt.flat_grad[17] = float("inf")
    opt.step()
    assert torch.equal(opt.flat_param, before) and opt.step_count() == 0          # GradScaler.step skipped the update
    assert float(opt.state[0]) == 32768.0 and opt.last_grad_norm() == float("inf")  # backoff 0.5
    assert opt.exp_avg.abs().sum().item() == 0
    opt.flat_grad.normal_()
    opt.flat_grad[5] = float("nan")
    opt.step()
    assert torch.equal(opt.flat_param, before) and float(opt.state[0]) == 16384.0
    for i in range(3):                                                            # three clean steps -> growth x2
        opt.flat_grad.normal_()
        opt.step()
    assert opt.step_count() == 3 and float(opt.state[0]) == 32768.0
    assert not torch.equal(opt.flat_param, before) and torch.isfinite(opt.flat_param).all()


def test_grad_norm_large_arena(cuda):
    from efficientsam3_b200 import ops
    n = 20_450_003                                                               # ~ TV-M parameter count, ragged tail
    g = torch.randn(n + 1, device=cuda)[:n] if False else torch.randn(n, device=cuda)
    ws, out = torch.zeros(8192, device=cuda), torch.zeros(2, device=cuda)
    ops.grad_norm(g, ws, out)
    ref = g.double().pow(2).sum().item()
    assert abs(out[0].item() - ref) <= 1e-6 * ref and out[1].item() == 0.0
    g[n - 1] = float("nan")
    ops.grad_norm(g, ws, out)
    assert out[1].item() == 1.0
