"""ctypes binding of include/adnb200.h.  The library is built in-tree by `adnm_unet_b200.build` (nvcc, sm_100a).
There is deliberately no fallback: a missing library or a non-CUDA tensor raises."""
import ctypes as C
import os

import torch

ADN_F32, ADN_BF16 = 0, 1
WT_MAX_LEVELS = 8
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libadnb200.so")
_lib = None

fp = C.POINTER(C.c_float)

MIXER_FIELDS = ("dt_bias", "A_log", "D", "scale", "shift", "alpha1", "alpha2", "in_proj_w",
                "conv_13_x1_w", "conv_31_x1_w", "conv_13_x2_w", "conv_31_x2_w",
                "conv_13_bc1_w", "conv_31_bc1_w", "conv_13_bc2_w", "conv_31_bc2_w",
                "conv2d_w", "norm_w", "norm_b", "conv2d_z_w", "out_proj_w")


class AdnShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "H", "W", "D", "Di", "P", "G", "N", "dtype", "flags")]


class AdnWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in MIXER_FIELDS]


class AdnWeightGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in MIXER_FIELDS]


class AdnFfnShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "H", "W", "D", "C4", "dtype")]


FFN_FIELDS = ("w_in", "b_in", "w_dw", "b_dw", "w_out", "b_out")


class AdnFfnWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in FFN_FIELDS]


class AdnConvShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "H", "W", "Cin", "Cout", "dtype")]


class WtShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "C", "H", "W", "k", "levels", "has_bias", "dtype")]


class WtWeights(C.Structure):
    _fields_ = [("base_conv_w", C.c_void_p), ("base_conv_b", C.c_void_p), ("base_scale_w", C.c_void_p),
                ("wavelet_conv_w", C.c_void_p * WT_MAX_LEVELS), ("wavelet_scale_w", C.c_void_p * WT_MAX_LEVELS)]


class WtWeightGrads(C.Structure):
    _fields_ = WtWeights._fields_


EXPORTS = {
    "adnssd_workspace_bytes": (C.c_int, [C.POINTER(AdnShape)] + [C.POINTER(C.c_size_t)] * 3),
    "adnssd_forward": (C.c_int, [C.POINTER(AdnShape), C.POINTER(AdnWeights)] + [C.c_void_p] * 5),
    "adnssd_backward": (C.c_int, [C.POINTER(AdnShape), C.POINTER(AdnWeights), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.POINTER(AdnWeightGrads), C.c_void_p, C.c_void_p]),
    "wtconv_workspace_bytes": (C.c_int, [C.POINTER(WtShape)] + [C.POINTER(C.c_size_t)] * 3),
    "wtconv_forward": (C.c_int, [C.POINTER(WtShape), C.POINTER(WtWeights)] + [C.c_void_p] * 5),
    "wtconv_backward": (C.c_int, [C.POINTER(WtShape), C.POINTER(WtWeights), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.POINTER(WtWeightGrads), C.c_void_p, C.c_void_p]),
    "adn_threshold_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.c_int32, C.c_float,
                                       C.c_void_p, C.c_void_p]),
    "adn_eval_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.POINTER(C.c_int32), C.c_int32, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "adn_sumsq_workspace_floats": (C.c_int, []),
    "adn_sumsq_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adn_adamw_flat": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_float] * 5 +
                       [C.c_int32, C.c_float, C.c_float, C.c_void_p]),
    "adn_rmsnorm_forward": (C.c_int, [C.c_void_p] * 6 + [C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "adn_rmsnorm_backward": (C.c_int, [C.c_void_p] * 10 + [C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "adn_residual_forward": (C.c_int, [C.c_void_p] * 6 + [C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "adn_residual_backward": (C.c_int, [C.c_void_p] * 12 + [C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "adn_ffn_workspace_bytes": (C.c_int, [C.POINTER(AdnFfnShape)] + [C.POINTER(C.c_size_t)] * 3),
    "adn_ffn_forward": (C.c_int, [C.POINTER(AdnFfnShape), C.POINTER(AdnFfnWeights)] + [C.c_void_p] * 5),
    "adn_ffn_backward": (C.c_int, [C.POINTER(AdnFfnShape), C.POINTER(AdnFfnWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(AdnFfnWeights), C.c_void_p, C.c_void_p]),
    "adn_linear_workspace_bytes": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "adn_linear_forward": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "adn_linear_backward": (C.c_int, [C.c_void_p] * 6 + [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "adn_conv3x3_path": (C.c_int, [C.POINTER(AdnConvShape)]),
    "adn_conv3x3_workspace_bytes": (C.c_int, [C.POINTER(AdnConvShape), C.POINTER(C.c_size_t)]),
    "adn_conv3x3_forward": (C.c_int, [C.POINTER(AdnConvShape)] + [C.c_void_p] * 7),
    "adn_conv3x3_backward": (C.c_int, [C.POINTER(AdnConvShape)] + [C.c_void_p] * 10),
    "adn_nchw_pack_forward": (C.c_int, [C.c_void_p] * 5 + [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "adn_nchw_pack_backward": (C.c_int, [C.c_void_p] * 10 + [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "adn_plane_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_int32, C.c_void_p]),
    "adn_plane_mix_forward": (C.c_int, [C.c_void_p] * 9 + [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "adn_plane_mix_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "adn_plane_mix_backward": (C.c_int, [C.c_void_p] * 14 + [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "adn_act_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "adn_act_backward": (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "adn_gconv4_forward": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 7 + [C.c_void_p]),
    "adn_gconv4_backward": (C.c_int, [C.c_void_p] * 6 + [C.c_int32] * 7 + [C.c_void_p]),
    "adn_sdpa_forward": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_float, C.c_int32, C.c_void_p]),
    "adn_sdpa_backward": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_float, C.c_int32, C.c_void_p]),
    "adn_last_error": (C.c_char_p, []),
    "adn_abi_version": (C.c_int, []),
    "adn_device_supported": (C.c_int, []),
    "adn_launch_count": (C.c_ulonglong, []),
    "adn_prof_enable": (C.c_int, [C.c_int]),
    "adn_prof_count": (C.c_int, []),
    "adn_prof_get": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]),
    "adn_phase_enable": (C.c_int, [C.c_int]),
    "adn_phase_read": (C.c_int, [C.POINTER(C.c_ulonglong)]),
    "adn_cta_times_read": (C.c_int, [C.POINTER(C.c_ulonglong)]),
    "adn_selftest_umma": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adn_selftest_umma_shift": (C.c_int, [C.c_int] * 6 + [C.c_void_p] * 5),
    "adnssd_kernel_family": (C.c_int, [C.POINTER(AdnShape)]),
    "adn_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "adn_selftest_gemm": (C.c_int, [C.c_int] * 6 + [C.c_void_p, C.c_longlong, C.c_longlong] * 4 +
                          [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                           C.c_void_p, C.c_void_p]),
    "adn_bench_umma": (C.c_int, [C.c_int] * 6 + [C.c_void_p] * 2),
}


def lib_path():
    return _LIB_PATH


def load():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} not found: run `python -m adnm_unet_b200.build` (nvcc, sm_100a). "
                              "adnm-unet_b200 has no CPU / PyTorch fallback.")
        lib = C.CDLL(_LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.adn_abi_version() != 1:
            raise ImportError("libadnb200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"adnb200 {what} failed (code {rc}): {load().adn_last_error().decode()}")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return ADN_F32
    if t.dtype == torch.bfloat16:
        return ADN_BF16
    raise RuntimeError(f"adnb200: unsupported activation dtype {t.dtype} (float32 or bfloat16)")


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"adnb200: `{name}` is on {t.device}; the sm_100a library is the only implementation "
                           "(no CPU fallback)")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """Raw cudaStream_t of torch's current stream on `device` (default: the current device)."""
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(idx))


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def on_device(device):
    """Context that makes `device` current for the library call; free when it already is."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(device)


def scratch(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def launch_count() -> int:
    return int(load().adn_launch_count())


class profile:
    """with _lib.profile() as p: ...; p.records -> [(kernel name, ms)], after synchronising the device."""

    def __enter__(self):
        load().adn_prof_enable(1)
        self.records = []
        return self

    def __exit__(self, *exc):
        lib = load()
        torch.cuda.synchronize()
        name, ms = C.c_char_p(), C.c_float()
        for i in range(lib.adn_prof_count()):
            check(lib.adn_prof_get(i, C.byref(name), C.byref(ms)), "adn_prof_get")
            self.records.append((name.value.decode(), ms.value))
        lib.adn_prof_enable(0)
        return False
