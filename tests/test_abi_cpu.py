"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol include/adnb200.h
declares, validates shapes on the host, and the host modules mirror the reference's state_dict layout.
No compute call is made (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from adnm_unet_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from adnm_unet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "adnb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(adn\w+|adnssd_\w+|wtconv_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.adn_abi_version() == 1


def test_struct_layout_matches_header():
    from adnm_unet_b200 import _lib
    assert C.sizeof(_lib.AdnShape) == 40
    assert C.sizeof(_lib.AdnWeights) == 21 * 8 == C.sizeof(_lib.AdnWeightGrads)
    assert C.sizeof(_lib.WtShape) == 32
    assert C.sizeof(_lib.WtWeights) == (3 + 2 * _lib.WT_MAX_LEVELS) * 8


def test_workspace_query_and_shape_validation(lib):
    from adnm_unet_b200 import _lib
    a, b, c = (C.c_size_t() for _ in range(3))
    ok = _lib.AdnShape(B=16, H=128, W=128, D=32, Di=64, P=4, G=2, N=16, dtype=_lib.ADN_BF16, flags=0)
    assert lib.adnssd_workspace_bytes(ok, a, b, c) == 0
    T, dip, CC = 16 * 128 * 128, 208, 192
    assert a.value >= T * (dip + 2 * CC) * 2 and b.value > 0 and c.value > 0
    for field, val, msg in (("G", 3, b"ngroups"), ("Di", 66, b"multiple of 4"), ("P", 5, b"headdim"),
                            ("N", 3, b"d_state"), ("dtype", 7, b"dtype"), ("B", 0, b"positive")):
        bad = _lib.AdnShape(B=16, H=128, W=128, D=32, Di=64, P=4, G=2, N=16, dtype=_lib.ADN_BF16, flags=0)
        setattr(bad, field, val)
        assert lib.adnssd_workspace_bytes(bad, a, b, c) != 0
        assert msg in lib.adn_last_error(), (field, lib.adn_last_error())


def test_mixer_module_mirrors_reference_state_dict():
    import adnm_unet_b200 as A
    from oracle import adnssd_oracle as AO
    torch.manual_seed(0)
    m = A.Mamba2(d_model=32, headdim=4, d_state=16, layer_idx=0, chunk_size=256, bimamba=True)
    sd = m.state_dict()
    assert list(sd) == list(AO.PARAM_NAMES)
    ref = AO.init_params(32, 4, 16)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ref.items()}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 16, 32), 4, 4)


def test_mixer_module_init_matches_reference_rng_stream():
    """Same seed -> same initial parameters as the reference constructor (only checked where /root/reference exists)."""
    import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference not mounted")
    import adnm_unet_b200 as A
    ref = ref_loader.load_reference()
    torch.manual_seed(123)
    a = ref.ADNssd.Mamba2(d_model=64, headdim=4, d_state=16).state_dict()
    torch.manual_seed(123)
    b = A.Mamba2(d_model=64, headdim=4, d_state=16).state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_wtconv_module_mirrors_reference_state_dict():
    import adnm_unet_b200 as A
    from oracle import wtconv_oracle as WO
    m = A.WTConv2d(8, 8, kernel_size=5, wt_levels=3)
    sd = m.state_dict()
    assert list(sd) == WO.param_names(3, True)
    ref = WO.init_params(8, 5, 3)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ref.items()}
    assert torch.allclose(sd["wt_filter"], ref["wt_filter"]) and not m.wt_filter.requires_grad


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "adnm-unet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
