"""CPU: the strict-mode graphs (efficientsam3_b200/strict.py) on torch statements of the strict ops (tests/emu_strict.py) reproduce
the reference fixtures -- the host logic (module walk, BN folding, LiteMLA channel layout, block-diagonal grouped conv, head) is
right before a GPU sees it."""
from types import SimpleNamespace as NS

import pytest
import torch

import emu_strict
from helpers import load_golden, sd_from_keys


def _student(name, img, embed, sd):
    from efficientsam3_b200.stage1.model import build_image_student_model
    cfg = NS(MODEL=NS(BACKBONE=name), DATA=NS(IMG_SIZE=img), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=embed))
    m = build_image_student_model(cfg)
    m.load_state_dict(sd)
    return m.eval()


@pytest.mark.parametrize("fixture,name", [("evm_160", "efficientvit_b1"), ("ev_b0_160", "efficientvit_b0"), ("ev_b2_192", "efficientvit_b2")])
def test_strict_student_graph_matches_reference_fixture(monkeypatch, fixture, name):
    from efficientsam3_b200 import ops, strict
    emu_strict.install(monkeypatch)
    g = load_golden(fixture)
    sd = sd_from_keys(g["keys"], int(g["seed_w"]))
    img, embed = int(g["img"]), int(g["embed"])
    m = _student(name, img, embed, sd)
    x = torch.randn(int(g["batch"]), 3, img, img, generator=torch.Generator().manual_seed(int(g["seed_x"])))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))     # the strict graph refuses CPU tensors
    with ops.strict_precision():
        assert ops.precision() == "strict"
        out = strict.student_forward(m, x)
    assert ops.precision() == "bf16"
    ref = torch.as_tensor(g["out"])
    rel = ((out.double() - ref.double()).norm() / ref.double().norm()).item()
    print(f"{fixture}: strict graph on CPU emulation vs reference fixture rel-L2 {rel:.3e}")
    assert out.shape == ref.shape and rel < 2e-5
