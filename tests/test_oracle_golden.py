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


@pytest.mark.parametrize("name", ["evm_160", "evm_64_quadratic"])
def test_efficientvit_oracle_matches_reference(name):
    from oracle import efficientvit as O
    g = _load(name)
    sd = _sd_from_keys(g["keys"], int(g["seed_w"]))
    img, embed, batch = int(g["img"]), int(g["embed"]), int(g["batch"])
    x = torch.randn(batch, 3, img, img, generator=torch.Generator().manual_seed(int(g["seed_x"])))
    with torch.no_grad():
        out, stages = O.image_student_encoder(sd, x, embed, "b1", return_stages=True)
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
