"""Shared test helpers (CPU + GPU)."""
import os

import numpy as np
import torch

from oracle.weights import fill_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def sd_from_keys(keys, seed):
    sd = {}
    for rec in keys:
        k, shp, dt = str(rec).split("|")
        shape = tuple(int(s) for s in shp.split(",")) if shp else ()
        sd[k] = torch.zeros(shape, dtype=getattr(torch, dt))
    sd = fill_state_dict(sd, seed)
    # complex buffers (ViT freqs_cis) are derived constants, not weights: the modules compute them
    return {k: v for k, v in sd.items() if not v.is_complex()}


def rel_l2(got, ref):
    got, ref = torch.as_tensor(got).double(), torch.as_tensor(ref).double()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


def max_err_over_scale(got, ref):
    got, ref = torch.as_tensor(got).double(), torch.as_tensor(ref).double()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-30)).item()


def cosine(got, ref):
    got, ref = torch.as_tensor(got).double().flatten(), torch.as_tensor(ref).double().flatten()
    return (got @ ref / (got.norm() * ref.norm() + 1e-30)).item()
