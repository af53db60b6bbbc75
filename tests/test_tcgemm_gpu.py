"""The general tcgen05 GEMM of the wide path (csrc/tcgemm.cuh) through its C-ABI self-test entry, against torch.matmul in
fp64 on the same bf16-rounded operands: both operand orientations, ragged M / N / K (zero-filled tails), two accumulating
segments, per-sample batches, split-K with fp32 atomics, the parity mask, alpha, and sub-matrix views (row pitch > width)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K0, K1=0, a_mn=0, b_mn=0, batches=1, splitk=1, c_mode=1, mask=0, alpha=None, pad=0, seed=0):
    from adnm_unet_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(seed)

    def operand(rows, K, mn):
        # stored [rows][K] (K-major) or [K][rows] (MN-major), inside a wider buffer when pad > 0 (row pitch > width)
        r8 = lambda n: (n + 7) // 8 * 8        # row pitches are whole 16-byte pieces
        shape = (batches, K, r8(rows) + pad) if mn else (batches, rows, r8(K) + pad)
        full = torch.randn(shape, device=dev, generator=g).bfloat16()
        view = full[:, :, :rows] if mn else full[:, :, :K]
        math = view.transpose(1, 2) if mn else view          # (batches, rows, K)
        return full, math.double(), shape[2], shape[1] * shape[2]

    A0, a0, lda0, abs0 = operand(M, K0, a_mn)
    B0, b0, ldb0, bbs0 = operand(N, K0, b_mn)
    ref = a0 @ b0.transpose(1, 2)
    if K1:
        A1, a1, lda1, abs1 = operand(M, K1, a_mn)
        B1, b1, ldb1, bbs1 = operand(N, K1, b_mn)
        ref = ref + a1 @ b1.transpose(1, 2)
    else:
        A1 = B1 = None
        lda1 = ldb1 = 8
        abs1 = bbs1 = 0
    al = None
    if alpha is not None:
        al = torch.tensor([alpha], device=dev)
        ref = ref * alpha
    if mask:
        mm = (torch.arange(M, device=dev)[:, None] ^ torch.arange(N, device=dev)[None, :]) & 1
        ref = ref * (mm == 0)
    ldc = N + pad
    Cbuf = torch.zeros(batches, M, ldc, device=dev, dtype=torch.bfloat16 if c_mode == 0 else torch.float32)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.adn_selftest_gemm(M, N, K0, K1, a_mn, b_mn, _lib.ptr(A0), lda0, abs0, _lib.ptr(B0), ldb0, bbs0,
                                     _lib.ptr(A1), lda1, abs1, _lib.ptr(B1), ldb1, bbs1, _lib.ptr(Cbuf), ldc, M * ldc, c_mode,
                                     batches, splitk, _lib.ptr(al), mask, _lib.ptr(status), _lib.stream_ptr()), "adn_selftest_gemm")
    torch.cuda.synchronize()
    assert int(status) == 0, "pipeline time-out flagged"
    got = Cbuf[:, :, :N].double()
    if pad:
        assert float(Cbuf[:, :, N:].abs().max()) == 0.0          # nothing written outside the matrix
    scale = float(ref.abs().max())
    return float((got - ref).abs().max()) / scale


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 208, 32), (200, 72, 136), (16, 256, 32), (640, 128, 1216), (2048, 32, 16)])
def test_orientations_and_ragged_shapes(a_mn, b_mn, M, N, K):
    assert _run(M, N, K, a_mn=a_mn, b_mn=b_mn) < 1e-5              # fp32 accumulation of exact bf16 products


def test_k_not_a_multiple_of_8_for_token_reductions():
    # reductions over tokens (MN-major operands): K = L is arbitrary (36 tokens of a 6 x 6 grid, 561 of a 33 x 17 grid)
    for K in (36, 561, 9):
        assert _run(256, 32, K, a_mn=1, b_mn=1) < 1e-5


def test_two_segments_batches_alpha_and_bf16_output():
    assert _run(144, 96, 64, K1=64, batches=3, alpha=0.37) < 1e-5
    assert _run(300, 128, 256, K1=256, c_mode=0, alpha=1.5, pad=8) < 4e-3      # bf16 rounding of the stored result
    assert _run(64, 256, 512, K1=512, a_mn=0, b_mn=1, batches=2) < 1e-5        # dy . S'^T with hi + lo


@pytest.mark.parametrize("splitk", [2, 5, 16])
def test_split_k_atomics_and_parity_mask(splitk):
    assert _run(256, 32, 4096, a_mn=1, b_mn=1, batches=2, splitk=splitk, c_mode=2, mask=1) < 1e-5
    assert _run(640, 128, 8192, a_mn=1, b_mn=1, splitk=splitk, c_mode=2, pad=16) < 1e-5


def test_padded_views_leave_neighbours_untouched():
    assert _run(130, 40, 72, pad=24) < 1e-5
    assert _run(130, 40, 72, a_mn=1, b_mn=1, pad=24, c_mode=0) < 4e-3
