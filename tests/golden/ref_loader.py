"""Import the UNMODIFIED reference (kanyu369/ADNM-UNet) for golden generation and the reference cross-checks.

The shims (stand-ins for the absent `timm` / `pywt` / `mamba_ssm`, the `.to('cuda')` no-op for CPU runs and the
size-generic `Decoder.forward`) live in `adnm_unet_b200.refhost`, because the full-model harness hosts the same
unmodified reference around the sm_100a drop-ins; this module only re-exports them under the names the golden
generator and the tests use.  No reference file is edited or copied into the tracked tree."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from adnm_unet_b200.refhost import (cuda_to_is_noop, load_reference, reference_available,  # noqa: E402,F401
                                    reference_root, install_shims as _install_shims)

REFERENCE_ROOT = reference_root()
