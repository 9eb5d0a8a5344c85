"""CPU: pin the oracle (oracle/*.py) against fixtures produced by the unmodified reference
(tests/golden/gen_golden.py).  Tolerance: fp32 round-off only (same math, different op order)."""
import os

import numpy as np
import pytest
import torch

from oracle.weights import fill_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


from helpers import sd_from_keys as _sd_from_keys


def _assert_close(got, ref, rtol=1e-5):
    scale = float(np.abs(ref).max())
    err = float(np.abs(got - ref).max())
    assert err <= rtol * scale + 1e-7, f"max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("name,variant", [("evm_160", "b1"), ("evm_64_quadratic", "b1"), ("ev_b0_160", "b0"), ("ev_b2_192", "b2")])
def test_efficientvit_oracle_matches_reference(name, variant):
    from oracle import efficientvit as O
    g = _load(name)
    sd = _sd_from_keys(g["keys"], int(g["seed_w"]))
    img, embed, batch = int(g["img"]), int(g["embed"]), int(g["batch"])
    x = torch.randn(batch, 3, img, img, generator=torch.Generator().manual_seed(int(g["seed_x"])))
    with torch.no_grad():
        out, stages = O.image_student_encoder(sd, x, embed, variant, return_stages=True)
    _assert_close(out.numpy(), g["out"], rtol=2e-5)
    for k, t in stages.items():
        assert tuple(t.shape) == tuple(g[f"shape_{k}"])
        ref = g[f"stats_{k}"]
        got = np.array([t.double().mean().item(), t.double().abs().mean().item(), t.double().std().item()])
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-6)
    n_params = sum(v.numel() for k, v in sd.items() if "running_" not in k and "num_batches" not in k)
    assert n_params == int(g["n_params"])


def test_vit_oracle_matches_reference():
    from oracle import vitdet as O
    g = _load("vit_small_112")
    cfg = eval(str(g["cfg"]))
    sd = _sd_from_keys(g["keys"], int(g["seed_w"]))
    x = torch.randn(int(g["batch"]), 3, cfg["img_size"], cfg["img_size"], generator=torch.Generator().manual_seed(int(g["seed_x"])))
    with torch.no_grad():
        out = O.vit_trunk(sd, "", x, cfg)
    _assert_close(out.numpy(), g["out"], rtol=2e-5)


def test_sam_heads_oracle_matches_reference():
    """PromptEncoder + TwoWayTransformer + MaskDecoder oracle vs the reference classes' recorded outputs."""
    import sys
    sys.path.insert(0, GOLD)
    from oracle import sam_heads as O
    g = _load("sam_heads_16")
    E, S, B = int(g["E"]), int(g["S"]), int(g["B"])
    sd_pe = _sd_from_keys(g["keys_pe"], int(g["seed_pe"]))
    sd_md = _sd_from_keys(g["keys_md"], int(g["seed_md"]))
    gen = torch.Generator().manual_seed(int(g["seed_x"]))
    feat = torch.randn(B, 256, E, E, generator=gen)
    f288 = torch.randn(B, 256, 4 * E, 4 * E, generator=gen)
    f144 = torch.randn(B, 256, 2 * E, 2 * E, generator=gen)
    coords = torch.rand(B, 1, 2, generator=gen) * S
    labels = torch.ones(B, 1, dtype=torch.int32)
    with torch.no_grad():
        sp, de = O.prompt_encoder_points(sd_pe, "", coords, labels, (S, S), (E, E))
        dpe = O.dense_pe(sd_pe, "", E, E)
        hr = O.high_res_from_fpn(sd_md, "", f288, f144)
        _assert_close(sp.numpy(), g["sparse"], rtol=1e-5)
        _assert_close(dpe.numpy()[:, :, ::4, ::4], g["dense_pe"], rtol=1e-5)
        for mm, sfx in ((True, "mm"), (False, "single")):
            m, iou, tok, obj = O.mask_decoder(sd_md, "", feat, dpe, sp, de, mm, hr)
            _assert_close(m.numpy(), g[f"masks_{sfx}"], rtol=5e-5)
            _assert_close(iou.numpy(), g[f"iou_{sfx}"], rtol=5e-5)
            _assert_close(tok.numpy(), g[f"tok_{sfx}"], rtol=5e-5)
            _assert_close(obj.numpy(), g[f"obj_{sfx}"], rtol=5e-5)
        q, k = O.two_way_transformer(sd_md, "transformer.", feat, dpe.expand(B, -1, -1, -1), torch.cat([sp, sp], dim=1))
        _assert_close(q.numpy(), g["twoway_q"], rtol=5e-5)


def test_neck_oracle_matches_reference():
    from oracle import necks as O
    g = _load("neck_small")
    sd = _sd_from_keys(g["keys"], int(g["seed_w"]))
    x = torch.randn(int(g["B"]), int(g["dim"]), int(g["hw"]), int(g["hw"]), generator=torch.Generator().manual_seed(int(g["seed_x"])))
    with torch.no_grad():
        for pref, name in (("convs.", "sam3"), ("sam2_convs.", "sam2")):
            outs = O.neck(sd, x, prefix=pref)
            for i, t in enumerate(outs):
                _assert_close(t.numpy(), g[f"{name}_{i}"], rtol=2e-5)


@pytest.mark.parametrize("name,variant", [("rvm_160", "repvit_m1_1"), ("rv_m0_9_128", "repvit_m0_9"), ("rv_m2_3_128", "repvit_m2_3")])
def test_repvit_oracle_matches_reference(name, variant):
    from oracle import repvit as O
    g = _load(name)
    sd = _sd_from_keys(g["keys"], int(g["seed_w"]))
    x = torch.randn(int(g["batch"]), 3, int(g["img"]), int(g["img"]), generator=torch.Generator().manual_seed(int(g["seed_x"])))
    with torch.no_grad():
        out = O.image_student_encoder(sd, x, int(g["embed"]), variant)
    _assert_close(out.numpy(), g["out"], rtol=2e-5)


@pytest.mark.parametrize("name,variant", [("tvm_160", "tiny_vit_11m"), ("tv_5m_160", "tiny_vit_5m"), ("tv_21m_160", "tiny_vit_21m")])
def test_tinyvit_oracle_matches_reference(name, variant):
    from oracle import tinyvit as O
    g = _load(name)
    sd = _sd_from_keys(g["keys"], int(g["seed_w"]))
    x = torch.randn(int(g["batch"]), 3, int(g["img"]), int(g["img"]), generator=torch.Generator().manual_seed(int(g["seed_x"])))
    with torch.no_grad():
        out = O.image_student_encoder(sd, x, int(g["embed"]), variant)
    _assert_close(out.numpy(), g["out"], rtol=2e-5)


def test_prompt_and_postprocess_oracle_matches_reference():
    """Box / mask prompts, repeat_image decoding and hole filling (+ resize) vs outputs of the reference classes."""
    import torch.nn.functional as F
    from oracle import sam_heads as O
    g = _load("sam_prompts_12")
    E, S, P = int(g["E"]), int(g["S"]), int(g["P"])
    sd_pe = _sd_from_keys(g["keys_pe"], int(g["seed_pe"]))
    sd_md = _sd_from_keys(g["keys_md"], int(g["seed_md"]))
    gen = torch.Generator().manual_seed(int(g["seed_x"]))
    feat = torch.randn(1, 256, E, E, generator=gen)
    f288 = torch.randn(1, 256, 4 * E, 4 * E, generator=gen)
    f144 = torch.randn(1, 256, 2 * E, 2 * E, generator=gen)
    coords = torch.rand(P, 2, 2, generator=gen) * S
    labels = torch.tensor([[1, 0], [1, 1], [0, 1]], dtype=torch.int32)
    xy0 = torch.rand(P, 2, generator=gen) * S * 0.5
    boxes = torch.cat([xy0, xy0 + 8 + torch.rand(P, 2, generator=gen) * S * 0.4], dim=1)
    mask_in = torch.randn(P, 1, 4 * E, 4 * E, generator=gen) * 4
    with torch.no_grad():
        sp, de = O.prompt_encoder(sd_pe, "", (coords, labels), boxes, mask_in, (S, S), (E, E))
        _assert_close(sp.numpy(), g["sparse_pts_boxes"], rtol=2e-5)
        _assert_close(de[:1].numpy(), g["dense_mask0"], rtol=2e-5)
        for i in range(P):
            got = np.array([de[i].double().mean().item(), de[i].double().abs().mean().item(), de[i].double().std().item()])
            np.testing.assert_allclose(got, g["dense_mask_stats"][i], rtol=1e-4, atol=1e-6)
        hr = O.high_res_from_fpn(sd_md, "", f288, f144)
        for mm, sfx in ((True, "mm"), (False, "single")):
            masks, iou, low = O.predict(sd_pe, sd_md, feat, hr, coords, labels, boxes, mask_in, S, (4 * E, 4 * E), multimask_output=mm,
                                        return_logits=True, max_hole_area=0.0)
            _assert_close(low.numpy(), np.clip(g[f"masks_{sfx}"], -32, 32), rtol=5e-5)
            _assert_close(iou.numpy(), g[f"iou_{sfx}"], rtol=5e-5)
        post = F.interpolate(O.fill_holes(torch.from_numpy(g["post_in"]), 0.0, 12.0, 5.0), (50, 70), mode="bilinear", align_corners=False)
    assert int(g["post_changed_px"]) > 100
    assert np.array_equal(post.numpy(), g["post_out"])       # same fp32 ops on the same labels: exact


def test_efficientvit_train_mode_oracle_matches_reference_training_step():
    """The oracle under bn_batch_stats() + oracle.kd_loss + autograd vs one training iteration of the unmodified reference
    student in .train() with the reference's own masked_mse / masked_cosine_loss / build_valid_mask (fixture in fp64:
    output, losses, every parameter gradient's norm / sum / leading entries, updated BN running statistics)."""
    from oracle import efficientvit as O
    from oracle.kd_loss import kd_loss
    g = _load("evm_train_160")
    img, embed, batch = int(g["img"]), int(g["embed"]), int(g["batch"])
    sd32 = _sd_from_keys(g["keys"], int(g["seed_w"]))
    sd = {k: ((v.double().requires_grad_(True) if "running" not in k else v.double()) if v.is_floating_point() else v.clone())
          for k, v in sd32.items()}
    gen = torch.Generator().manual_seed(int(g["seed_x"]))
    x = torch.randn(batch, 3, img, img, generator=gen).double()
    teacher = torch.randn(batch, 1024, embed, embed, generator=gen).double()
    sizes = [tuple(int(v) for v in r) for r in g["sizes"]]
    with O.bn_batch_stats():
        out = O.image_student_encoder(sd, x, embed, "b1")
    loss, mse, cos = kd_loss(out, teacher, img, sizes, float(g["cosine"]))
    loss.backward()
    np.testing.assert_allclose(out.detach()[:, ::8].numpy(), g["out"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose([loss.item(), mse.item(), cos.item()], g["loss"], rtol=1e-10)
    scale = float(g["grad_stats"][:, 0].max())
    for name, ref in zip(g["grad_names"], g["grad_stats"]):
        gr = sd[str(name)].grad.reshape(-1)
        got = np.concatenate([[gr.norm().item(), gr.sum().item()], np.pad(gr[:4].numpy(), (0, max(0, 4 - gr.numel())))])
        np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-9 * scale, err_msg=str(name))
    for i, name in enumerate(g["buf_names"]):
        np.testing.assert_allclose(sd[str(name)].numpy(), g[f"buf{i}"], rtol=1e-10, atol=1e-12, err_msg=str(name))


def test_student_vision_backbone_oracle_matches_reference():
    """EfficientSAM3 image encoder (student trunk + 1024-channel head + dual FPN) as `_create_student_vision_backbone` builds it:
    oracle composition (efficientvit + student head + neck) vs the reference's outputs on both FPN branches."""
    from oracle import efficientvit as EV
    from oracle import necks as NK
    g = _load("student_neck_evm_160")
    sd = _sd_from_keys(g["keys"], int(g["seed_w"]))
    img = int(g["img"])
    x = torch.randn(1, 3, img, img, generator=torch.Generator().manual_seed(int(g["seed_x"])))
    student = {k[len("trunk.model."):]: v for k, v in sd.items() if k.startswith("trunk.model.")}
    with torch.no_grad():
        feats = EV.image_student_encoder(student, x, 72, "b1")
        assert tuple(feats.shape) == (1, 1024, 72, 72)
        for name, prefix in (("sam3", "convs."), ("sam2", "sam2_convs.")):
            for i, t in enumerate(NK.neck(sd, feats, prefix=prefix)):
                assert tuple(t.shape) == tuple(g[f"{name}_{i}_shape"])
                st = 6 if t.shape[-1] > 36 else 3
                _assert_close(t[:, ::16, ::st, ::st].numpy(), g[f"{name}_{i}_sub"], rtol=5e-5)


def test_student_vision_backbone_keys_match_reference_for_all_nine_students():
    """Key-for-key (name, shape, dtype, order) equality of efficientsam3_b200.model_builder.create_student_vision_backbone with the
    reference builder, through the sha256 of the signature list recorded from the reference."""
    import hashlib
    from efficientsam3_b200.model_builder import create_student_vision_backbone
    g = _load("student_neck_evm_160")
    want = dict(zip([str(k) for k in g["sig_names"]], [str(v) for v in g["sig_values"]]))
    assert len(want) == 9
    for key, ref in want.items():
        bt, mn = key.split(":")
        sd = create_student_vision_backbone(bt, mn, enable_inst_interactivity=True).state_dict()
        ks = [f"{k}|{','.join(map(str, v.shape))}|{str(v.dtype).replace('torch.', '')}" for k, v in sd.items()]
        assert f"{len(ks)}:{hashlib.sha256(chr(10).join(ks).encode()).hexdigest()}" == ref, key
    with pytest.raises(ValueError):
        create_student_vision_backbone("resnet", "50")
