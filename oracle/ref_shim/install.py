"""Third-party shim that makes the UNMODIFIED reference importable in this container.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference (/root/reference) depends on
packages that are not installed here (timm>=1.0.17 per sam3/pyproject.toml:18, iopath, ftfy,
omegaconf, pycocotools, decord, yacs, mmengine, termcolor).  Only a handful of symbols are
actually reached on the hot path; they are restated below from their published semantics:

* timm.layers.DropPath      -- per-sample Bernoulli keep, scaled by 1/keep_prob, identity in eval
* timm.layers.Mlp           -- fc1 -> act -> drop1 -> norm(Identity) -> fc2 -> drop2
                               (call site: sam3/sam3/model/vitdet.py:585-590)
* timm.layers.SqueezeExcite -- x * sigmoid(fc2(relu(fc1(mean_HW(x))))), 1x1 convs WITH bias,
                               rd_channels = make_divisible(C*rd_ratio, 8, round_limit=0.)
                               (call site: sam3/sam3/backbones/repvit.py:23,136,150)
* timm.layers.trunc_normal_ / to_2tuple, timm.models.register_model,
  timm.models._builder.build_model_with_cfg (constructs the class; no pretrained download)

Nothing from /root/reference is copied.  `install()` registers the stub modules in sys.modules
and puts the reference's package roots on sys.path.
"""
from __future__ import annotations

import collections.abc
import math
import sys
import types
from itertools import repeat

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


# ----------------------------------------------------------------------------- timm restatements
def _ntuple(n):
    def parse(x):
        if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
            return tuple(x)
        return tuple(repeat(x, n))

    return parse


to_2tuple = _ntuple(2)


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def drop_path(x, drop_prob: float = 0.0, training: bool = False, scale_by_keep: bool = True):
    if drop_prob == 0.0 or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    random_tensor = x.new_empty(shape).bernoulli_(keep_prob)
    if keep_prob > 0.0 and scale_by_keep:
        random_tensor.div_(keep_prob)
    return x * random_tensor


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training, self.scale_by_keep)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0.0, use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        bias = to_2tuple(bias)
        drop_probs = to_2tuple(drop)
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0])
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop_probs[0])
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1])
        self.drop2 = nn.Dropout(drop_probs[1])

    def forward(self, x):
        x = self.fc1(x)
        x = self.act(x)
        x = self.drop1(x)
        x = self.norm(x)
        x = self.fc2(x)
        x = self.drop2(x)
        return x


def make_divisible(v, divisor=8, min_value=None, round_limit=0.9):
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


class SqueezeExcite(nn.Module):
    """timm.layers.squeeze_excite.SEModule (aliased SqueezeExcite)."""

    def __init__(self, channels, rd_ratio=1.0 / 16, rd_channels=None, rd_divisor=8, add_maxpool=False,
                 bias=True, act_layer=nn.ReLU, norm_layer=None, gate_layer=nn.Sigmoid):
        super().__init__()
        self.add_maxpool = add_maxpool
        if not rd_channels:
            rd_channels = make_divisible(channels * rd_ratio, rd_divisor, round_limit=0.0)
        self.fc1 = nn.Conv2d(channels, rd_channels, kernel_size=1, bias=bias)
        self.bn = norm_layer(rd_channels) if norm_layer else nn.Identity()
        self.act = act_layer(inplace=True)
        self.fc2 = nn.Conv2d(rd_channels, channels, kernel_size=1, bias=bias)
        self.gate = gate_layer()

    def forward(self, x):
        x_se = x.mean((2, 3), keepdim=True)
        if self.add_maxpool:
            x_se = 0.5 * x_se + 0.5 * x.amax((2, 3), keepdim=True)
        x_se = self.fc1(x_se)
        x_se = self.act(self.bn(x_se))
        x_se = self.fc2(x_se)
        return x * self.gate(x_se)


def register_model(fn):
    return fn


def build_model_with_cfg(model_cls, variant, pretrained, pretrained_cfg=None, default_cfg=None,
                         pretrained_filter_fn=None, **kwargs):
    assert not pretrained, "shim: no pretrained weights offline"
    return model_cls(**kwargs)


# ----------------------------------------------------------------------------- module plumbing
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so submodule imports resolve through sys.modules
    sys.modules[name] = m
    return m


class _PathMgr:
    def open(self, path, mode="r", **kw):
        return open(path, mode)

    def exists(self, path):
        import os
        return os.path.exists(path)

    def __getattr__(self, name):  # anything else is off the hot path
        raise AttributeError(f"shim g_pathmgr has no {name}")


_installed = False


def install(reference_root: str = REFERENCE_ROOT):
    global _installed
    if _installed:
        return
    layers = dict(DropPath=DropPath, Mlp=Mlp, trunc_normal_=trunc_normal_, to_2tuple=to_2tuple,
                  SqueezeExcite=SqueezeExcite, drop_path=drop_path)
    if "timm" not in sys.modules:
        timm = _mod("timm", __version__="1.0.17")
        timm.layers = _mod("timm.layers", **layers)
        timm.models = _mod("timm.models", register_model=register_model)
        timm.models.layers = _mod("timm.models.layers", **layers)
        timm.models.vision_transformer = _mod("timm.models.vision_transformer", trunc_normal_=trunc_normal_)
        timm.models._builder = _mod("timm.models._builder", build_model_with_cfg=build_model_with_cfg)
        timm.models.helpers = _mod("timm.models.helpers", build_model_with_cfg=build_model_with_cfg)
        timm.models.registry = _mod("timm.models.registry", register_model=register_model)

    def _try(name):
        try:
            __import__(name)
            return True
        except Exception:
            return False

    if not _try("iopath"):
        iopath = _mod("iopath")
        iopath.common = _mod("iopath.common")
        iopath.common.file_io = _mod("iopath.common.file_io", g_pathmgr=_PathMgr())
    if not _try("ftfy"):
        _mod("ftfy", fix_text=lambda s: s)
    if not _try("omegaconf"):
        class _OC:
            @staticmethod
            def create(*a, **k):
                raise RuntimeError("shim OmegaConf")
        _mod("omegaconf", MISSING="???", OmegaConf=_OC, DictConfig=dict, ListConfig=list)
    if not _try("pycocotools"):
        pc = _mod("pycocotools")
        pc.mask = _mod("pycocotools.mask")
    if not _try("decord"):
        _mod("decord")
    if not _try("skimage"):
        # scikit-image is absent here; the reference's CPU connected-components backend (perflib/connected_components.py:
        # 18-29) calls skimage.measure.label(values, return_num=True): default connectivity = input.ndim, i.e.
        # 8-connected components of the non-zero pixels in 2-D, numbered from 1.  Restated with scipy.ndimage.label.
        def _label(values, return_num=False, connectivity=None):
            import numpy as np
            from scipy import ndimage
            lab, n = ndimage.label(np.asarray(values) != 0, structure=np.ones((3,) * np.ndim(values), dtype=np.int32))
            return (lab.astype(np.int64), n) if return_num else lab.astype(np.int64)
        sk = _mod("skimage")
        sk.measure = _mod("skimage.measure", label=_label)
    if not _try("termcolor"):
        _mod("termcolor", colored=lambda s, *a, **k: s)

    # `import sam3` must resolve to the OUTER package (/root/reference/sam3/__init__.py), whose
    # __path__ shim also exposes the inner one; stage1/ is a flat script directory.
    for p in (f"{reference_root}/stage1", reference_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True
