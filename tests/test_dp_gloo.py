"""N>1 host logic on CPU: world-size-2 gloo, batch sharding + gradient all-reduce == single-process big-batch gradient.
The per-rank compute is the CPU oracle (tests may use it); the product's DP plumbing (adnm_unet_b200.dp) is what is tested."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import adnssd_oracle as AO

D, P, N, G, GLOBAL_B = 16, 4, 8, 6, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grads(params, u, dout, numel):
    q = {k: v.clone().requires_grad_(k not in AO.UNUSED_PARAMS) for k, v in params.items()}
    out = AO.mixer_forward(q, u, G, G, P, N)
    ((out * dout).sum() / numel).backward()
    return q


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adnm_unet_b200.dp import GradAllReducer, shard_range
    torch.manual_seed(0)
    params = AO.init_params(D, P, N, seed=3, perturb=0.2, dtype=torch.float64)
    g = torch.Generator().manual_seed(7)
    u = torch.randn(GLOBAL_B, G * G, D, generator=g, dtype=torch.float64)
    dout = torch.randn(GLOBAL_B, G * G, D, generator=g, dtype=torch.float64)
    a, b = shard_range(GLOBAL_B, rank, world)
    q = _grads(params, u[a:b], dout[a:b], u[a:b].numel())
    plist = list(q.values())
    red = GradAllReducer(plist)
    red()
    norm = red.grad_norm().item()
    if rank == 0:
        ref = _grads(params, u, dout, u.numel())
        errs = {}
        for k in q:
            if ref[k].grad is None:
                assert q[k].grad is None, k            # unused parameters stay grad-less on every rank
            else:
                errs[k] = ((q[k].grad - ref[k].grad).abs().max() / ref[k].grad.abs().max().clamp_min(1e-300)).item()
        ref_norm = torch.cat([v.grad.reshape(-1) for v in ref.values() if v.grad is not None]).norm().item()
        ret["errs"], ret["norm"], ret["ref_norm"] = errs, norm, ref_norm
    dist.destroy_process_group()


def test_sharded_grads_equal_big_batch_grads():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["errs"] and max(ret["errs"].values()) < 1e-10, dict(ret["errs"])
    assert abs(ret["norm"] - ret["ref_norm"]) / ret["ref_norm"] < 1e-10


def test_shard_range_rejects_ragged_batches():
    from adnm_unet_b200.dp import shard_range
    assert shard_range(32, 3, 8) == (12, 16)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
