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


STUDENTS = [("evm_160", "efficientvit_b1"), ("ev_b0_160", "efficientvit_b0"), ("ev_b2_192", "efficientvit_b2"), ("rvm_160", "repvit_m1_1"),
            ("rv_m0_9_128", "repvit_m0_9"), ("rv_m2_3_128", "repvit_m2_3"), ("tvm_160", "tiny_vit_11m"), ("tv_5m_160", "tiny_vit_5m"),
            ("tv_21m_160", "tiny_vit_21m")]


@pytest.mark.parametrize("fixture,name", STUDENTS)
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


def test_strict_vit_graph_matches_reference_fixture(monkeypatch):
    from efficientsam3_b200 import ops
    from efficientsam3_b200.model.vitdet import create_sam3_vit_backbone
    emu_strict.install(monkeypatch)
    monkeypatch.setattr(ops, "tokens_f32_to_nchw", lambda xs, B, h, w: xs.view(B, h, w, -1).permute(0, 3, 1, 2).contiguous())
    g = load_golden("vit_small_112")
    cfg = eval(str(g["cfg"]))
    m = create_sam3_vit_backbone(**cfg)
    m.load_state_dict(sd_from_keys(g["keys"], int(g["seed_w"])), strict=False)
    m.eval()
    x = torch.randn(int(g["batch"]), 3, cfg["img_size"], cfg["img_size"], generator=torch.Generator().manual_seed(int(g["seed_x"])))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    with ops.strict_precision():
        out = m(x)[-1]
    ref = torch.as_tensor(g["out"])
    rel = ((out.double() - ref.double()).norm() / ref.double().norm()).item()
    print(f"vit_small_112: strict graph on CPU emulation vs reference fixture rel-L2 {rel:.3e}")
    assert out.shape == ref.shape and rel < 2e-5
