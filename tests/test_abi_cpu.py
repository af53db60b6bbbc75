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


def test_block_level_entry_points_validate_on_the_host(lib):
    """The Block / attention / evaluator entry points reject bad shapes and dtypes before touching the device."""
    from adnm_unet_b200 import _lib
    assert C.sizeof(_lib.AdnFfnShape) == 24 and C.sizeof(_lib.AdnFfnWeights) == 6 * 8
    a, b, c = (C.c_size_t() for _ in range(3))
    ok = _lib.AdnFfnShape(B=32, H=128, W=128, D=32, C4=128, dtype=_lib.ADN_BF16)
    assert lib.adn_ffn_workspace_bytes(ok, a, b, c) == 0
    T = 32 * 128 * 128
    assert a.value >= T * (128 + 128 + 64) * 2 and b.value > 0 and c.value >= T * (64 + 128 + 128) * 2
    for field, val in (("D", 30), ("C4", 100), ("dtype", 5), ("B", 0)):
        bad = _lib.AdnFfnShape(B=32, H=128, W=128, D=32, C4=128, dtype=_lib.ADN_BF16)
        setattr(bad, field, val)
        assert lib.adn_ffn_workspace_bytes(bad, a, b, c) != 0 and lib.adn_last_error()
    one = C.c_void_p(256)      # never dereferenced: validation comes first
    assert lib.adn_sdpa_forward(one, one, None, 2, 16, 8, 5, 0.5, _lib.ADN_BF16, None) != 0 and b"dim_head" in lib.adn_last_error()
    assert lib.adn_sdpa_forward(one, one, None, 2, 16, 8, 4, 0.5, 9, None) != 0 and b"dtype" in lib.adn_last_error()
    assert lib.adn_rmsnorm_forward(one, one, None, None, one, None, 10, 30, 1e-6, _lib.ADN_F32, None) != 0
    assert lib.adn_residual_forward(one, one, one, one, None, one, 0, 32, _lib.ADN_F32, None) != 0
    nb = C.c_size_t()
    assert lib.adn_linear_workspace_bytes(100, 64, 32, _lib.ADN_BF16, nb) == 0 and nb.value >= 64 * 32 * 2
    assert lib.adn_linear_workspace_bytes(0, 64, 32, _lib.ADN_BF16, nb) != 0
    thr = (C.c_int32 * 4)(20, 30, 35, 40)
    assert lib.adn_eval_batch(one, one, 4000, 20, 16, thr, 4, 90.0, one, one, None) != 0      # batch * seq_len > 65535


def test_block_modules_mirror_reference_state_dict():
    """Block / FeedForward / RMSNorm / StandardAttention: reference key names, shapes and registration order; same-seed
    construction consumes the RNG stream like the reference classes (only checked where the reference is importable)."""
    import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference not mounted")
    from adnm_unet_b200 import refhost
    for dim, out_dim in ((32, 32), (64, 128)):
        a = refhost.build_block(dim, out_dim, dropin=False, seed=11).state_dict()
        b = refhost.build_block(dim, out_dim, dropin=True, seed=11).state_dict()
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k
    import adnm_unet_b200 as A
    ref = ref_loader.load_reference()
    torch.manual_seed(5)
    ra = ref.ADNssd.StandardAttention(64, heads=16, dim_head=4, dropout=0.).state_dict()
    torch.manual_seed(5)
    rb = A.StandardAttention(64, heads=16, dim_head=4, dropout=0.).state_dict()
    assert list(ra) == list(rb) and all(torch.equal(ra[k], rb[k]) for k in ra)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        A.RMSNorm(16)(torch.zeros(2, 3, 16))


def test_convstage_modules_mirror_reference_state_dict():
    """WTLayer / PatchEmbed / OutProj: reference key names, shapes and registration order; a same-seed
    construction consumes the RNG stream like the reference classes; the hosted network with the stages bound has the
    reference's state_dict, key for key and bit for bit (only checked where the reference is importable)."""
    import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference not mounted")
    import cases
    from adnm_unet_b200 import convstage, refhost
    ref = ref_loader.load_reference()
    for name, (kind, kw, B, g, skip) in cases.CONVSTAGE_CASES.items():
        torch.manual_seed(17)
        a = getattr(ref.model_untils, kind)(**kw)
        torch.manual_seed(17)
        b = getattr(convstage, kind)(**kw)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb), name
        assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()], name
        for k in sa:      # the frozen Haar filters: exact +-0.5 here, float32 (1/sqrt 2)^2 = 0.49999997 from the pywt taps there
            assert torch.allclose(sa[k], sb[k], rtol=0, atol=1e-7) if k.endswith("wt_filter") else torch.equal(sa[k], sb[k]), (name, k)
        assert torch.equal(torch.rand(3), (torch.manual_seed(17), getattr(ref.model_untils, kind)(**kw), torch.rand(3))[2])
    a = refhost.build_adnm_unet(128, dropin=False, seed=0).state_dict()
    m = refhost.build_adnm_unet(128, dropin=True, seed=0)
    b = m.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.allclose(a[k], b[k], rtol=0, atol=1e-7) if k.endswith("wt_filter") else torch.equal(a[k], b[k]), k
    assert type(m.encoder.encoder1) is convstage.PatchEmbed and type(m.decoder.decoder6) is convstage.WTLayer
    assert type(m.refiner.out_proj) is convstage.OutProj and type(m.encoder.encoder2) is convstage.WTLayer
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.encoder.encoder2(torch.zeros(1, 64, 32))


def test_convstage_entry_points_validate_on_the_host(lib):
    from adnm_unet_b200 import _lib
    assert C.sizeof(_lib.AdnConvShape) == 24
    nb = C.c_size_t()
    ok = _lib.AdnConvShape(B=32, H=128, W=128, Cin=64, Cout=32, dtype=_lib.ADN_BF16)
    assert lib.adn_conv3x3_workspace_bytes(ok, nb) == 0 and nb.value >= 32 * 9 * 64 * (2 + 4) + 64 * 9 * 64 * 2
    assert lib.adn_conv3x3_path(ok) == 1                                    # tcgen05 implicit GEMM
    assert lib.adn_conv3x3_path(_lib.AdnConvShape(B=32, H=128, W=128, Cin=5, Cout=32, dtype=_lib.ADN_BF16)) == 1     # thin: padded rows
    thin = _lib.AdnConvShape(B=32, H=128, W=128, Cin=20, Cout=20, dtype=_lib.ADN_BF16)
    assert lib.adn_conv3x3_workspace_bytes(thin, nb) == 0 and nb.value >= 3 * 32 * 128 * 128 * 24 * 2
    assert lib.adn_conv3x3_path(_lib.AdnConvShape(B=2, H=24, W=24, Cin=64, Cout=32, dtype=_lib.ADN_BF16)) == 0      # 24 does not tile
    assert lib.adn_conv3x3_path(_lib.AdnConvShape(B=2, H=16, W=16, Cin=64, Cout=32, dtype=_lib.ADN_F32)) == 0       # check mode
    assert lib.adn_conv3x3_path(_lib.AdnConvShape(B=4, H=4, W=4, Cin=128, Cout=256, dtype=_lib.ADN_BF16)) == 1       # boxes span samples
    for field, val in (("B", 0), ("Cin", 0), ("dtype", 3)):
        bad = _lib.AdnConvShape(B=32, H=128, W=128, Cin=64, Cout=32, dtype=_lib.ADN_BF16)
        setattr(bad, field, val)
        assert lib.adn_conv3x3_workspace_bytes(bad, nb) != 0 and lib.adn_last_error()
    one = C.c_void_p(256)      # never dereferenced: validation comes first
    assert lib.adn_plane_stats(one, one, 0, 16, 1e-5, _lib.ADN_F32, None) != 0
    assert lib.adn_plane_mix_forward(one, one, None, None, None, one, one, None, one, 2, 8, 64, 2, _lib.ADN_F32, None) != 0 and b"act" in lib.adn_last_error()
    assert lib.adn_act_forward(one, one, 10, 3, None, _lib.ADN_F32, None) != 0 and b"kind" in lib.adn_last_error()
    assert lib.adn_nchw_pack_forward(one, None, None, None, one, 2, 64, 8, 8, _lib.ADN_F32, None) != 0      # C2 > 0 without res
    assert lib.adn_plane_mix_workspace_bytes(32, 64, nb) == 0 and nb.value >= 32 * 64 * 16
    assert lib.adn_gconv4_forward(one, one, None, one, 2, 4, 4, 30, 3, 3, _lib.ADN_BF16, None) != 0 and b"multiple of 4" in lib.adn_last_error()
    assert lib.adn_gconv4_forward(one, one, None, one, 2, 4, 4, 32, 5, 3, _lib.ADN_BF16, None) != 0 and b"kernel" in lib.adn_last_error()
