"""Drop-in for the reference `stage1/model.py` (image path): same builder / class names, same
state_dict keys, forward on libes3.so.

  build_image_student_model(config)   stage1/model.py:30-39
  ImageStudentEncoder                 stage1/model.py:188-211   (head.0/1/3 keys preserved)
  EfficientViTAdapter                 stage1/model.py:327-335
  _build_backbone                     stage1/model.py:386-417

`config` needs MODEL.BACKBONE, DATA.IMG_SIZE, DISTILL.EMBED_DIM, DISTILL.EMBED_SIZE (yacs CfgNode or any
attribute namespace).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..backbones.efficientvit import (efficientvit_backbone_b0, efficientvit_backbone_b1,
                                      efficientvit_backbone_b2)
from ..nn_utils import NativePlanMixin, bn_scale_bias, conv3x3_weight, params_fingerprint, pw_weight, pw_weight_scaled


def build_image_student_model(config):
    backbone_name = config.MODEL.BACKBONE.lower()
    backbone, out_channels = _build_backbone(backbone_name, config.DATA.IMG_SIZE)
    return ImageStudentEncoder(backbone=backbone, in_channels=out_channels, embed_dim=config.DISTILL.EMBED_DIM,
                               embed_size=config.DISTILL.EMBED_SIZE, img_size=config.DATA.IMG_SIZE)


class ImageStudentEncoder(nn.Module, NativePlanMixin):
    def __init__(self, backbone, in_channels, embed_dim, embed_size, img_size):
        super().__init__()
        self.backbone = backbone
        self.embed_size = embed_size
        self.img_size = img_size
        # parameter containers; keys head.0.weight, head.1.{weight,bias,running_*}, head.3.{weight,bias}
        self.head = nn.Sequential(
            nn.Conv2d(in_channels, embed_dim, kernel_size=1, bias=False),
            nn.BatchNorm2d(embed_dim),
            nn.GELU(),
            nn.Conv2d(embed_dim, embed_dim, kernel_size=3, padding=1),
        )

    def _build_plan(self):
        dev = self.head[0].weight.device
        s, b = bn_scale_bias(self.head[1], None, self.head[0].out_channels, dev)
        return dict(w0=pw_weight_scaled(self.head[0], s), b0=b, w3=conv3x3_weight(self.head[3]),
                    b3=self.head[3].bias.detach().float().contiguous())

    def forward(self, x):
        if self.training:
            return self._forward_train(x)
        with torch.no_grad():
            if ops.precision() == "strict":           # fp32 activations / weights / accumulation (strict.py, csrc/strict_f32.cu)
                from ..strict import student_forward
                return student_forward(self, x)
            if self._graphs is not None and x.is_cuda:
                return self._forward_graphed(x)
            return self._forward_eval(x)

    # ---- CUDA-graph replay of the eval plan ----------------------------------------------------------------------------------
    # The eval forward is a fixed sequence of ~60 kernels whose arguments depend only on (input address, shape, packed weights).
    # Launching it kernel by kernel costs ~1 ms of host time per step, which is invisible on one GPU (the device needs ~5 ms) but
    # becomes the limiter when 8 ranks share one host (round-1 SCALE: 0.934 efficiency with no collective in the step).  With
    # `enable_cuda_graphs()` the sequence is captured once per (input buffer, shape) and replayed with ONE launch.
    _graphs = None
    graph_launches_per_step = 0

    def enable_cuda_graphs(self, enabled: bool = True, max_graphs: int = 4):
        """Replay the eval forward from a CUDA graph.  A graph is bound to the ADDRESS of its input tensor (feed a small rotating set
        of input buffers, as a double-buffered loader does) and to the current parameter values; it is re-captured when either moves.
        The returned tensor is the graph's own output buffer: it is overwritten by the next replay for the same input buffer."""
        self._graphs = {} if enabled else None
        self._graph_max = max_graphs
        return self

    def forward_uncaptured(self, x):
        with torch.no_grad():
            return self._forward_eval(x)

    def _forward_graphed(self, x):
        key = (x.data_ptr(), tuple(x.shape), x.dtype)
        fp = params_fingerprint(self)
        ent = self._graphs.get(key)
        if ent is None or ent["fp"] != fp:
            self._forward_eval(x)                      # un-captured pass: packs weights, sizes workspaces, configures kernels
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count
            with torch.cuda.graph(graph):              # private memory pool per graph: outputs of different graphs never alias
                out = self._forward_eval(x)
            ent = dict(graph=graph, out=out, fp=fp, x=x, launches=ops.launch_count - n0)   # holds x: the graph reads its address
            self._graphs.pop(key, None)
            while len(self._graphs) >= self._graph_max:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = ent
            self.graph_launches_per_step = ent["launches"]
        ent["graph"].replay()
        return ent["out"]

    def _forward_train(self, x):
        """Train-mode forward recorded as ONE autograd node (StudentTrainFunction): batch-statistics BatchNorm (or frozen
        BN modules, set_bn_state), backward on the kernels of train_bwd.cu.  Built for all nine students (EfficientViT b0 / b1 / b2, RepViT m0_9 / m1_1 / m2_3, TinyViT 5m / 11m / 21m)."""
        if not isinstance(self.backbone, (EfficientViTAdapter, RepViTAdapter, TinyViTAdapter)):
            raise NotImplementedError(
                "train-mode forward / backward is built for the nine students of the reference builder (EfficientViT, RepViT, "
                f"TinyViT adapters); {type(self.backbone).__name__} is eval-only.  Call .eval() first.")
        params = [p for p in self.parameters()]
        return StudentTrainFunction.apply(self, x, *params)

    def _forward_eval(self, x):
        if x.dtype in (torch.bfloat16, torch.float16):
            # half-width image batches (half the host->device bytes).  The first kernel of every student rounds the image to bf16
            # operands anyway (tensor-core stem), so a bf16 batch gives the results of its fp32 original for the fused EV-M stem.
            x = x.float()
        feats = self.backbone.forward_nhwc(x)          # [B,h,w,Cin] bf16
        p = self._plan()
        B, h, w, cin = feats.shape
        y = ops.gemm(feats.view(-1, cin), p["w0"], bias=p["b0"], act="gelu")       # BN scale folded into w0 (nn_utils.pw_weight_scaled)
        y = ops.conv3x3(y.view(B, h, w, -1), p["w3"], bias=p["b3"])
        if h != self.embed_size or w != self.embed_size:
            return ops.bilinear_nhwc_to_nchw(y, self.embed_size, self.embed_size)
        return ops.nhwc_to_nchw_f32(y)


class StudentTrainFunction(torch.autograd.Function):
    """preds = model(samples) with a native backward: forward runs the training graph (backbones/efficientvit_train.py) and
    keeps it; backward(d preds) walks it in reverse and returns one fp32 gradient per parameter, so `loss.backward()`,
    `p.grad`, DDP's gradient hooks and GradScaler work exactly as with the reference nn.Module."""

    @staticmethod
    def forward(ctx, module, x, *params):
        from ..backbones.efficientvit_train import EfficientViTTrainGraph, HeadTrainUnit
        ctx.module = module
        from ..backbones.repvit_train import RepViTTrainGraph
        from ..backbones.tinyvit_train import TinyViTTrainGraph
        if not (x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
            raise ValueError("expected an fp32 NCHW image batch [B,3,H,W]")
        for m in module.modules():          # packed eval-mode weights go stale once parameters / running stats move
            if isinstance(m, NativePlanMixin):
                m._plan_key = None
        if isinstance(module.backbone, RepViTAdapter):
            body = RepViTTrainGraph(module.backbone.model)
        elif isinstance(module.backbone, TinyViTAdapter):
            body = TinyViTTrainGraph(module.backbone.model)
        else:
            body = EfficientViTTrainGraph(module.backbone.model)
        head = HeadTrainUnit(module.head, module.embed_size)
        out = head.forward(body.forward(x))
        ctx.graph = (body, head)
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.graph is None:
            raise RuntimeError("the native student keeps its saved activations for ONE backward pass (retain_graph / double backward "
                               "are not supported): run the forward again")
        body, head = ctx.graph
        ctx.graph = None
        from ..backbones.efficientvit_train import GradSink
        grads = GradSink()
        arena = getattr(ctx.module, "_es3_grad_arena", None)       # set by stage1.optim.FlatAdamW(direct_grads=True)
        grads.direct = arena is not None
        d = head.backward(dout, grads)
        if arena is not None:
            arena.head_grads_ready()                               # the head's weights (2/3 of EV-M's parameters) are final: exchange them now
        body.backward(d, grads)
        return (None, None) + tuple(grads.get(p) for p in ctx.params)


def build_image_teacher_model(config):
    """stage1/model.py:168-175.  `config.MODEL.RESUME` (optional) is a reference SAM3 checkpoint."""
    checkpoint = getattr(config.MODEL, "RESUME", None) or None
    teacher = SAM3ImageTeacherEncoder(checkpoint_path=checkpoint, embed_size=config.DISTILL.EMBED_SIZE)
    teacher.img_size = config.DATA.IMG_SIZE
    return teacher


class _Holder(nn.Module):
    """Attribute container that reproduces the reference's key prefix `sam3.backbone.vision_backbone.trunk.`"""


class SAM3ImageTeacherEncoder(nn.Module):
    """stage1/model.py:214-249: frozen SAM3 ViT trunk -> [B,1024,72,72].  Only the trunk is instantiated (the
    reference builds the whole SAM3 model and then uses nothing else on this path); a full reference checkpoint
    loads with strict=False through the preserved key prefix."""

    def __init__(self, checkpoint_path=None, embed_size=64, vit_overrides=None):
        super().__init__()
        from ..model.vitdet import create_sam3_vit_backbone
        self.embed_size = embed_size
        self.sam3 = _Holder()
        self.sam3.backbone = _Holder()
        self.sam3.backbone.vision_backbone = _Holder()
        self.sam3.backbone.vision_backbone.trunk = create_sam3_vit_backbone(**(vit_overrides or {}))
        if checkpoint_path:
            sd = torch.load(checkpoint_path, map_location="cpu")
            sd = sd.get("model", sd)
            # reference checkpoints prefix the image model with "detector." (model_builder.py:584-630)
            sd = {("sam3." + k[len("detector."):] if k.startswith("detector.") else k): v for k, v in sd.items()}
            own = set(self.state_dict().keys())
            self.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        for p in self.parameters():
            p.requires_grad = False
        self.eval()
        self.img_size = 1008

    def train(self, mode: bool = True):  # frozen teacher: always eval (stage1/model.py:226-227)
        return super().train(False)

    @torch.no_grad()
    def forward(self, x):
        feats = self.sam3.backbone.vision_backbone.trunk(x)[-1]
        if feats.shape[-1] != self.embed_size or feats.shape[-2] != self.embed_size:
            # stage1/model.py:241-248 (bilinear, align_corners=False); a no-op in the shipped configs (EMBED_SIZE 72)
            feats, _ = ops.bilinear_nchw(feats, self.embed_size, self.embed_size)
        return feats


class RepViTAdapter(nn.Module):
    """stage1/model.py:287-296."""

    def __init__(self, model, out_channels):
        super().__init__()
        self.model = model
        self.out_channels = out_channels

    def forward(self, x):
        return ops.nhwc_to_nchw_f32(self.model.forward_nhwc(x))

    def forward_nhwc(self, x):
        return self.model.forward_nhwc(x)


class TinyViTAdapter(nn.Module):
    """stage1/model.py:299-324 (head / norm_head replaced by Identity)."""

    def __init__(self, model, img_size):
        super().__init__()
        self.model = model
        self.out_channels = self.model.norm_head.normalized_shape[0]
        self.model.head = nn.Identity()
        self.model.norm_head = nn.Identity()
        H, W = self.model.patches_resolution
        for _ in range(self.model.num_layers - 1):
            H, W = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        self.final_hw = (H, W)

    def forward(self, x):
        return ops.nhwc_to_nchw_f32(self.model.forward_nhwc(x))

    def forward_nhwc(self, x):
        return self.model.forward_nhwc(x)


class EfficientViTAdapter(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model
        self.out_channels = self.model.width_list[-1]

    def forward(self, x):
        return self.model(x)["stage_final"]

    def forward_nhwc(self, x):
        return self.model.forward_nhwc(x)


def _build_backbone(name, img_size):
    if name.startswith("efficientvit"):
        fn = {"efficientvit_b0": efficientvit_backbone_b0, "efficientvit_b1": efficientvit_backbone_b1,
              "efficientvit_b2": efficientvit_backbone_b2}[name]
        adapter = EfficientViTAdapter(fn())
        return adapter, adapter.out_channels
    if name in ("repvit_m0_9", "repvit_m1_1", "repvit_m2_3"):
        from ..backbones import repvit
        model = getattr(repvit, name)(pretrained=False, num_classes=0, distillation=False)
        out_channels = repvit._make_divisible(model.cfgs[-1][2], 8)
        return RepViTAdapter(model, out_channels), out_channels
    if name in ("tiny_vit_5m", "tiny_vit_11m", "tiny_vit_21m"):
        from ..backbones import tiny_vit
        adapter = TinyViTAdapter(getattr(tiny_vit, name + "_224")(pretrained=False, img_size=img_size), img_size)
        return adapter, adapter.out_channels
    raise ValueError(f"Unsupported backbone {name}")
