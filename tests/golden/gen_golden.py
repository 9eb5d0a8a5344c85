"""Generate golden fixtures by running the UNMODIFIED reference modules (CPU fp32) in the build container.

    python tests/golden/gen_golden.py [evm|teacher|decoder|all]

Needs /root/reference (absent on the GPU box -- fixtures are committed).  Weights come from
oracle.weights.fill_state_dict(seed) so they are reproducible without being stored.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))

import install  # noqa: E402  (oracle/ref_shim/install.py)

install.install()
from oracle.weights import fill_state_dict  # noqa: E402

torch.set_grad_enabled(False)


def keyshapes(sd):
    return np.array([f"{k}|{','.join(map(str, v.shape))}|{str(v.dtype).replace('torch.', '')}" for k, v in sd.items()])


def stats(t):
    t = t.double()
    return np.array([t.mean().item(), t.abs().mean().item(), t.std().item()])


def gen_student(backbone, tag, img, embed, seed_w, seed_x, batch=1):
    import model as stage1_model  # /root/reference/stage1/model.py

    cfg = NS(MODEL=NS(BACKBONE=backbone), DATA=NS(IMG_SIZE=img), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=embed))
    m = stage1_model.build_image_student_model(cfg).eval()
    sd = fill_state_dict(m.state_dict(), seed_w)
    m.load_state_dict(sd)
    x = torch.randn(batch, 3, img, img, generator=torch.Generator().manual_seed(seed_x))
    out = m(x)
    extra = {}
    if backbone.startswith("efficientvit"):
        st = m.backbone.model(x)
        for k in ("stage0", "stage1", "stage2", "stage3", "stage4"):
            extra[f"stats_{k}"] = stats(st[k])
            extra[f"shape_{k}"] = np.array(st[k].shape)
    path = os.path.join(HERE, f"{tag}.npz")
    np.savez_compressed(path, out=out.numpy(), keys=keyshapes(m.state_dict()), img=img, embed=embed, seed_w=seed_w,
                        seed_x=seed_x, batch=batch, n_params=sum(p.numel() for p in m.parameters()), **extra)
    print(tag, "out", tuple(out.shape), "absmean %.4f" % out.abs().mean().item(), "->", path,
          "%.1f KB" % (os.path.getsize(path) / 1024))


VIT_SMALL = dict(img_size=112, pretrain_img_size=56, patch_size=14, embed_dim=256, depth=4, num_heads=4, mlp_ratio=4.625,
                 window_size=4, global_att_blocks=(1, 3))


def gen_vit(tag, cfg, seed_w, seed_x, batch):
    """Reference ViT class (sam3/sam3/model/vitdet.py) in the SAM3 configuration family, small dims."""
    from sam3.model.vitdet import ViT

    m = ViT(norm_layer="LayerNorm", drop_path_rate=0.1, qkv_bias=True, use_abs_pos=True, tile_abs_pos=True,
            rel_pos_blocks=(), use_rope=True, use_interp_rope=True, pretrain_use_cls_token=True, retain_cls_token=False,
            ln_pre=True, ln_post=False, return_interm_layers=False, bias_patch_embed=False, **cfg).eval()
    sd = fill_state_dict(m.state_dict(), seed_w)
    m.load_state_dict(sd)
    x = torch.randn(batch, 3, cfg["img_size"], cfg["img_size"], generator=torch.Generator().manual_seed(seed_x))
    out = m(x)[-1]
    path = os.path.join(HERE, f"{tag}.npz")
    np.savez_compressed(path, out=out.numpy(), keys=keyshapes(m.state_dict()), seed_w=seed_w, seed_x=seed_x, batch=batch,
                        n_params=sum(p.numel() for p in m.parameters()), cfg=np.array(repr(cfg)))
    print(tag, "out", tuple(out.shape), "absmean %.4f" % out.abs().mean().item(), "->", path,
          "%.1f KB" % (os.path.getsize(path) / 1024))


def main(which):
    if which in ("vit", "all"):
        gen_vit("vit_small_112", VIT_SMALL, seed_w=21, seed_x=3, batch=2)
    if which in ("evm", "all"):
        # 160x160: stage3 10x10, stage4 5x5 (> dim=16 pixels so the linear-attention branch runs, odd size
        # exercises the stride-2 padding), head 5x5 -> bilinear to 12x12.
        gen_student("efficientvit_b1", "evm_160", img=160, embed=12, seed_w=11, seed_x=12, batch=1)
        # 64x64: stage4 is 2x2 = 4 <= dim -> the quadratic-attention branch (oracle only).
        gen_student("efficientvit_b1", "evm_64_quadratic", img=64, embed=2, seed_w=13, seed_x=14, batch=2)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "all")
