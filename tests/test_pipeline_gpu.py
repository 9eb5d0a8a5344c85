"""GPU: FPN neck vs the reference fixture, and the batched point-prompt pipeline (ViT trunk -> SAM2-branch FPN ->
SAM heads, BASELINE config 3 geometry at reduced ViT depth) against the composition of the CPU oracles."""
import pytest
import torch
import torch.nn.functional as F

from helpers import load_golden, max_err_over_scale, rel_l2, sd_from_keys

pytestmark = pytest.mark.gpu


class _Trunk(torch.nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.channel_list = [dim]

    def forward(self, x):
        return [x]


def test_neck_matches_reference_fixture(cuda):
    from efficientsam3_b200.model.necks import Sam3DualViTDetNeck
    g = load_golden("neck_small")
    dim, d_model, hw, B = int(g["dim"]), int(g["d_model"]), int(g["hw"]), int(g["B"])
    m = Sam3DualViTDetNeck(trunk=_Trunk(dim), position_encoding=None, d_model=d_model, scale_factors=[4.0, 2.0, 1.0, 0.5],
                           add_sam2_neck=True)
    sig = [f"{k}|{','.join(map(str, v.shape))}|{str(v.dtype).replace('torch.', '')}" for k, v in m.state_dict().items()]
    assert sig == [str(k) for k in g["keys"]]
    m.load_state_dict(sd_from_keys(g["keys"], int(g["seed_w"])))
    m = m.to(cuda).eval()
    x = torch.randn(B, dim, hw, hw, generator=torch.Generator().manual_seed(int(g["seed_x"]))).to(cuda)
    s3, _, s2, _ = m(x)
    for name, outs in (("sam3", s3), ("sam2", s2)):
        for i, t in enumerate(outs):
            ref = g[f"{name}_{i}"]
            assert t.shape == tuple(ref.shape)
            e = rel_l2(t.cpu(), ref)
            print(f"neck {name} level {i}: rel_l2={e:.3e}")
            assert e <= 1e-2


def test_point_prompt_pipeline_vs_oracles(cuda):
    from efficientsam3_b200.model.sam1_task import Sam3PointPromptSegmenter
    from oracle import necks as ON, sam_heads as OH, vitdet as OV
    from oracle.weights import fill_state_dict
    vit_cfg = dict(depth=2, global_att_blocks=(1,))
    seg = Sam3PointPromptSegmenter(vit_overrides=vit_cfg)
    sd = {k: v for k, v in fill_state_dict(seg.state_dict(), 41).items() if not v.is_complex()}
    seg.load_state_dict(sd, strict=False)
    B, S = 2, 1008
    g = torch.Generator().manual_seed(5)
    img = torch.randn(B, 3, S, S, generator=g)
    coords = torch.rand(B, 1, 2, generator=g) * S
    labels = torch.ones(B, 1, dtype=torch.int32)
    # ---- oracle composition on CPU
    with torch.no_grad():
        cfg = dict(OV.SAM3_VIT, **vit_cfg)
        trunk = OV.vit_trunk(sd, "backbone.vision_backbone.trunk.", img, cfg)
        l288, l144, l72 = ON.neck({k[len("backbone.vision_backbone."):]: v for k, v in sd.items() if k.startswith("backbone.vision_backbone.")},
                                  trunk, prefix="sam2_convs.")[:3]
        sd_md = {k[len("sam_mask_decoder."):]: v for k, v in sd.items() if k.startswith("sam_mask_decoder.")}
        sd_pe = {k[len("sam_prompt_encoder."):]: v for k, v in sd.items() if k.startswith("sam_prompt_encoder.")}
        hr = OH.high_res_from_fpn(sd_md, "", l288, l144)
        feat = l72 + sd["no_mem_embed"].reshape(1, -1, 1, 1)
        ref = OH.forward_sam_heads(sd_pe, sd_md, feat, hr, coords, labels, S, multimask_output=True)
    # ---- native
    seg = seg.to(cuda)
    out = seg.set_image_batch(img.to(cuda)).predict_batch(coords.to(cuda), labels.to(cuda), multimask_output=True, return_logits=True)
    low, high = out["low_res_multimasks"].cpu(), out["high_res"].cpu()
    e_low, e_obj = rel_l2(low, ref["low_res_multimasks"]), (out["object_score_logits"].cpu() - ref["object_score_logits"]).abs().max().item()
    print(f"pipeline low-res logits rel_l2={e_low:.3e}, obj abs err={e_obj:.3e}, ious err={(out['ious'].cpu() - ref['ious']).abs().max().item():.3e}")
    assert torch.equal(out["object_score_logits"].cpu() > 0, ref["object_score_logits"] > 0)
    assert e_low <= 2e-2 and e_obj <= 5e-2
    assert torch.equal(out["best"].cpu(), ref["best"])
    err = (high.double() - ref["high_res_multimasks"].double()).abs().max().item()
    safe = ref["high_res_multimasks"].abs() > err
    assert torch.equal((high > 0)[safe], (ref["high_res_multimasks"] > 0)[safe])
    agree = ((high > 0) == (ref["high_res_multimasks"] > 0)).float().mean().item()
    print(f"binary mask agreement {agree:.5f}, rounding band {(~safe).float().mean().item():.4%}")
    assert agree >= 0.99
    bm = seg.predict_batch(coords.to(cuda), labels.to(cuda), multimask_output=True)["high_res"]
    assert bm.dtype == torch.bool and torch.equal(bm.cpu(), high > 0)
