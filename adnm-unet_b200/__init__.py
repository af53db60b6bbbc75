"""adnm-unet_b200: B200-native (sm_100a) ADN-SSD token mixer, Haar WTConv2d and threshold counts behind the
reference's own module API (kanyu369/ADNM-UNet: models/ADNssd.py::Mamba2, models/WTConv2d.py::WTConv2d,
datasets/Shanghai_metrics.py).  All compute goes through the C ABI in include/adnb200.h (ctypes); there is no
CPU or PyTorch fallback - calling an op without the built CUDA library or without a GPU raises."""
from adnm_unet_b200 import _lib  # noqa: F401
from adnm_unet_b200.mixer import Mamba2, adnssd_mixer  # noqa: F401
from adnm_unet_b200.wtconv import WTConv2d, wtconv2d  # noqa: F401
from adnm_unet_b200.metrics import threshold_counts, csi_hss  # noqa: F401
from adnm_unet_b200.rmsnorm import RMSNorm, rmsnorm_affine  # noqa: F401
from adnm_unet_b200.block import Block, FeedForward, residual_mix, linear_tokens, make_block  # noqa: F401
from adnm_unet_b200.attention import StandardAttention, sdpa_packed  # noqa: F401
from adnm_unet_b200.evaluator import SimplifiedEvaluator  # noqa: F401
from adnm_unet_b200.convstage import WTLayer, PatchEmbed, OutProj, conv3x3_tokens, plane_mix, pack_planes  # noqa: F401
from adnm_unet_b200.inject import install_into_reference  # noqa: F401
from adnm_unet_b200.dp import GradAllReducer, shard_range  # noqa: F401
from adnm_unet_b200.graphed import graphed_mixer  # noqa: F401

__all__ = ["Mamba2", "adnssd_mixer", "WTConv2d", "wtconv2d", "threshold_counts", "csi_hss", "install_into_reference",
           "graphed_mixer", "RMSNorm", "rmsnorm_affine", "Block", "FeedForward", "residual_mix", "linear_tokens", "make_block",
           "StandardAttention", "sdpa_packed", "SimplifiedEvaluator", "WTLayer", "PatchEmbed", "OutProj", "conv3x3_tokens", "plane_mix",
           "pack_planes"]
