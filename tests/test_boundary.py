"""CPU: the C-ABI boundary.  The library loads, exports every function include/es3.h declares, the ctypes
table binds them with the declared arity, and the product package never touches oracle/."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    """name -> number of parameters, parsed from include/es3.h."""
    src = open(os.path.join(ROOT, "include", "es3.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long long|const char\*)\s+(es3_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
    return out


def test_header_symbols_are_exported_and_bound():
    from efficientsam3_b200 import _lib
    lib = _lib.load()
    decl = _declared()
    assert len(decl) >= 20, decl
    for name, nargs in decl.items():
        fn = getattr(lib, name)  # raises AttributeError if the .so does not export it
        assert isinstance(fn, ctypes._CFuncPtr)
        if name in _lib.SIGNATURES:
            assert len(_lib.SIGNATURES[name]) == nargs, (name, nargs, len(_lib.SIGNATURES[name]))
    missing = [n for n in _lib.SIGNATURES if n not in decl]
    assert not missing, f"bound but not declared in include/es3.h: {missing}"
    assert lib.es3_version() >= 100


def test_no_gpu_is_a_loud_error_not_a_fallback():
    import pytest
    import torch
    from efficientsam3_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.Es3Error):
        _lib.init(0)
    with pytest.raises(_lib.Es3Error):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(32, 8, dtype=torch.bfloat16))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "efficientsam3_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dp, f)
                assert "/root/reference" not in text, os.path.join(dp, f)


STUDENT_FIXTURES = {"efficientvit_b0": "ev_b0_160", "efficientvit_b1": "evm_160", "efficientvit_b2": "ev_b2_192",
                    "repvit_m0_9": "rv_m0_9_128", "repvit_m1_1": "rvm_160", "repvit_m2_3": "rv_m2_3_128",
                    "tiny_vit_5m": "tv_5m_160", "tiny_vit_11m": "tvm_160", "tiny_vit_21m": "tv_21m_160"}


def test_unknown_backbone_is_a_value_error():
    from types import SimpleNamespace as NS
    import pytest
    from efficientsam3_b200.stage1.model import build_image_student_model
    with pytest.raises(ValueError):
        build_image_student_model(NS(MODEL=NS(BACKBONE="resnet50"), DATA=NS(IMG_SIZE=160), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=12)))


def test_state_dict_keys_match_reference_record():
    """Key-for-key (name, shape, dtype, order) equality with the key list recorded from the reference modules."""
    from types import SimpleNamespace as NS
    from helpers import load_golden
    from efficientsam3_b200.stage1.model import build_image_student_model
    from efficientsam3_b200.model.vitdet import create_sam3_vit_backbone

    def sig(sd):
        return [f"{k}|{','.join(map(str, v.shape))}|{str(v.dtype).replace('torch.', '')}" for k, v in sd.items()]

    # all nine names the reference builder accepts (stage1/model.py:386-417)
    for name, fixture in STUDENT_FIXTURES.items():
        g = load_golden(fixture)
        cfg = NS(MODEL=NS(BACKBONE=name), DATA=NS(IMG_SIZE=int(g["img"])), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=int(g["embed"])))
        m = build_image_student_model(cfg)
        assert sig(m.state_dict()) == [str(k) for k in g["keys"]], name
        assert sum(p.numel() for p in m.parameters()) == int(g["n_params"]), name
    g = load_golden("vit_small_112")
    assert sig(create_sam3_vit_backbone(**eval(str(g["cfg"]))).state_dict()) == [str(k) for k in g["keys"]]
