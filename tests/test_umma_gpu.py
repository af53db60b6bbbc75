"""tcgen05 building blocks (shared-memory descriptors in the T8 layout, instruction descriptor, TMEM read-back)
checked against torch.matmul on exactly representable bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1], ids=["k_major", "mn_major"])
@pytest.mark.parametrize("N,K", [(16, 16), (32, 32), (64, 64), (208, 32), (128, 128), (256, 64), (48, 256)])
def test_umma_selftest(mode, N, K):
    from adnm_unet_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    if mode == 0:
        A = torch.randint(-8, 9, (128, K), generator=g).float()
        B = torch.randint(-8, 9, (N, K), generator=g).float()
        ref = A @ B.t()
    else:
        A = torch.randint(-8, 9, (K, 128), generator=g).float()
        B = torch.randint(-8, 9, (K, N), generator=g).float()
        ref = A.t() @ B
    Ad, Bd = A.cuda().bfloat16().contiguous(), B.cuda().bfloat16().contiguous()
    C = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.adn_selftest_umma(mode, N, K, _lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(C), _lib.ptr(status),
                                     _lib.stream_ptr()), "adn_selftest_umma")
    torch.cuda.synchronize()
    assert status.item() == 0, "MMA completion barrier timed out"
    assert torch.equal(C.cpu(), ref), (C.cpu() - ref).abs().max()


@pytest.mark.parametrize("mode", [0, 1], ids=["k_major", "mn_major"])
@pytest.mark.parametrize("N,K,pitch,sa,sb", [(32, 32, 130, 0, 0), (32, 32, 130, 1, 0), (32, 32, 130, 2, 1), (96, 48, 129, 1, 0),
                                             (32, 128, 130, 0, 2), (208, 32, 258, 129, 3), (64, 64, 137, 7, 5)])
def test_umma_shift(mode, N, K, pitch, sa, sb):
    """Row-padded operand buffers addressed with row-shifted (16-byte granular) descriptor start addresses."""
    from adnm_unet_b200 import _lib
    lib = _lib.load()
    if mode == 1 and (sa + K > pitch or sb + K > pitch):
        pytest.skip("extent")
    g = torch.Generator().manual_seed(N * 1000 + K + mode + pitch)
    if mode == 0:
        A = torch.randint(-8, 9, (pitch, K), generator=g).float()
        B = torch.randint(-8, 9, (pitch, K), generator=g).float()
        ref = A[sa:sa + 128] @ B[sb:sb + N].t()
    else:
        A = torch.randint(-8, 9, (pitch, 128), generator=g).float()
        B = torch.randint(-8, 9, (pitch, N), generator=g).float()
        ref = A[sa:sa + K].t() @ B[sb:sb + K]
    Ad, Bd = A.cuda().bfloat16().contiguous(), B.cuda().bfloat16().contiguous()
    C = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.adn_selftest_umma_shift(mode, N, K, pitch, sa, sb, _lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(C),
                                           _lib.ptr(status), _lib.stream_ptr()), "adn_selftest_umma_shift")
    torch.cuda.synchronize()
    assert status.item() == 0, "MMA completion barrier timed out"
    assert torch.equal(C.cpu(), ref), (C.cpu() - ref).abs().max()
