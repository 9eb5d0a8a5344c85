"""GPU: stage-1 KD loss kernel vs the oracle restatement of train_image_encoder_stage1.py:271-307 (fp32: rtol 1e-4)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,E,img", [(4, 1024, 72, 1008), (3, 64, 16, 224), (2, 1024, 64, 1024)])
def test_kd_loss(cuda, B, C, E, img):
    from efficientsam3_b200.stage1.losses import kd_loss
    from oracle import kd_loss as O
    g = torch.Generator().manual_seed(B * E)
    p = torch.randn(B, C, E, E, generator=g)
    t = (p * 0.7 + 0.5 * torch.randn(B, C, E, E, generator=g)).half().float()
    # even samples padded on the right (w = 3/4), odd ones at the bottom (h = 2/3); one full image
    sizes = [(3, img, img * 3 // 4) if i % 2 == 0 else (3, img * 2 // 3, img) for i in range(B)]
    sizes[-1] = (3, img, img)
    ref = O.kd_loss(p, t, img, sizes, cosine_weight=1.0)
    got = kd_loss(p.to(cuda), t.to(cuda), img, sizes, cosine_weight=1.0)
    for a, b, n in zip(got, ref, ("loss", "mse", "cos")):
        assert abs(a.item() - b.item()) <= 1e-4 * abs(b.item()) + 1e-6, (n, a.item(), b.item())
