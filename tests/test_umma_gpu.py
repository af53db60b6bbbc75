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
