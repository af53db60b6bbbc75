"""Host logic of the batch-sharded trainer (adnm_unet_b200.trainer) on CPU: world-size-2 gloo.

Two ranks, each with half of the global batch, must end every step with the SAME parameters as one process running
`clip_grad_norm_` + `torch.optim.AdamW` + `zero_grad` on the whole batch (train.py:132-146 semantics).  The per-step tail
is injected as a plain-torch restatement (the product default is the CUDA library and refuses CPU tensors); what is under
test is discovery of the live parameter set, flattening, bucketing, the overlapped all-reduce hooks and the 1/world scale.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Net(nn.Module):
    """Small network with the two features that matter: a parameter that never gets a gradient and a shared one."""

    def __init__(self):
        super().__init__()
        self.a = nn.Linear(6, 16)
        self.b = nn.Linear(16, 16)
        self.c = nn.Linear(16, 3)
        self.unused = nn.Parameter(torch.ones(5))
        self.beta = nn.Parameter(torch.tensor(0.7))

    def forward(self, x):
        h = torch.tanh(self.a(x))
        h = self.beta * h + self.beta * torch.tanh(self.b(h))
        return self.c(h)


def torch_tail(p, g, m, v, ws, step, lr, grad_scale, clip_norm, hp):
    """Plain-torch restatement of adn_sumsq_f32 + adn_adamw_flat (csrc/optim.cu)."""
    with torch.no_grad():
        norm = g.norm() * grad_scale
        coef = grad_scale * (min(1.0, clip_norm / (norm.item() + 1e-6)) if clip_norm else 1.0)
        ge = g * coef
        m.mul_(hp["beta1"]).add_(ge, alpha=1 - hp["beta1"])
        v.mul_(hp["beta2"]).addcmul_(ge, ge, value=1 - hp["beta2"])
        p.mul_(1 - lr * hp["weight_decay"])
        denom = v.sqrt() / (1 - hp["beta2"] ** step) ** 0.5 + hp["eps"]
        p.addcdiv_(m, denom, value=-lr / (1 - hp["beta1"] ** step))
        g.zero_()
        ws["norm"].fill_(norm)


def _loss(out, tgt):
    return (out - tgt).abs().sum() / tgt.numel()


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adnm_unet_b200.trainer import DataParallelTrainer, shard_range
    from adnm_unet_b200.refhost import ADAMW
    torch.manual_seed(0)
    net = Net()
    ref = Net()
    ref.load_state_dict(net.state_dict())
    hp = dict(ADAMW)
    tr = DataParallelTrainer(net, _loss, clip_norm=0.05, adamw=hp, bucket_bytes=512, autocast_dtype=None,
                             step_tail=torch_tail)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(3, 8, 6, generator=g, dtype=torch.float32)
    Y = torch.randn(3, 8, 3, generator=g, dtype=torch.float32)
    opt = torch.optim.AdamW(ref.parameters(), lr=hp["lr"], betas=(hp["beta1"], hp["beta2"]), eps=hp["eps"],
                            weight_decay=hp["weight_decay"])
    a, b = shard_range(8, rank, world)
    norms = []
    for s in range(3):
        tr.step(X[s, a:b], Y[s, a:b])
        norms.append(float(tr.grad_norm()))
        _loss(ref(X[s]), Y[s]).backward()
        rn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        opt.step()
        opt.zero_grad()
        norms[-1] = (norms[-1], float(rn))
    err = max(float((p - q).abs().max()) for p, q in zip(net.parameters(), ref.parameters()))
    if rank == 0:
        ret["err"], ret["norms"] = err, norms
        ret["unused_untouched"] = bool(torch.equal(net.unused, torch.ones(5))) and net.unused.grad is None
        ret["buckets"], ret["live"] = len(tr.buckets), len(tr.live)
    dist.destroy_process_group()


def test_two_ranks_match_single_process_big_batch_training():
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["err"] < 2e-6, ret["err"]
    for mine, theirs in ret["norms"]:
        assert abs(mine - theirs) / theirs < 1e-5
    assert ret["unused_untouched"] and ret["live"] == 7 and ret["buckets"] >= 2


def test_reference_lr_schedule_matches_torch_sequential_lr():
    """train_untils.py:44-46"""
    from adnm_unet_b200.trainer import reference_lr
    p = [nn.Parameter(torch.zeros(1))]
    opt = torch.optim.AdamW(p, lr=1e-3)
    w = torch.optim.lr_scheduler.LinearLR(opt, start_factor=0.01, total_iters=3)
    c = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50, eta_min=5e-7)
    sch = torch.optim.lr_scheduler.SequentialLR(opt, [w, c], [3])
    for epoch in range(12):
        assert abs(opt.param_groups[0]["lr"] - reference_lr(epoch)) < 1e-9, epoch
        opt.step()
        sch.step()


def test_grad_allreducer_keeps_aliased_grads():
    """a second call with the previous views still installed must not copy a slice onto itself."""
    from adnm_unet_b200.dp import GradAllReducer
    ps = [nn.Parameter(torch.randn(4, 3)), nn.Parameter(torch.randn(5))]
    for p in ps:
        p.grad = torch.ones_like(p)
    red = GradAllReducer(ps)
    red()
    for p in ps:
        p.grad.add_(1.0)          # accumulate into the views (set_to_none=False style)
    red()
    assert all(bool((p.grad == 2).all()) for p in ps)


def test_dead_bridge_pruning_is_bit_identical_on_cpu():
    """refhost.prune_dead_bridges: e2ds[3..6] never reach the output (models/ADNMUNet.py:603-630, model_untils.py:407-408);
    skipping them leaves output, every gradient and the set of grad-less parameters bit-identical (reference modules, CPU)."""
    import torch
    from adnm_unet_b200 import refhost
    if not refhost.reference_available():
        import pytest
        pytest.skip("reference sources not available")
    a = refhost.build_adnm_unet(128, dropin=False, seed=0)
    b = refhost.build_adnm_unet(128, dropin=False, seed=0, prune_dead=True)
    assert list(a.state_dict()) == list(b.state_dict())
    x = torch.rand(1, 5, 1, 128, 128, generator=torch.Generator().manual_seed(3))
    with refhost.cuda_to_is_noop():
        ya, yb = a(x), b(x)
        ya.square().sum().backward()
        yb.square().sum().backward()
    assert torch.equal(ya, yb)
    ga, gb = ({k: p.grad for k, p in m.named_parameters()} for m in (a, b))
    assert all((ga[k] is None) == (gb[k] is None) for k in ga)
    assert all(torch.equal(ga[k], gb[k]) for k in ga if ga[k] is not None)


def test_graph_mode_is_eager_on_cpu_tensors():
    """DataParallelTrainer(graph=True) only captures CUDA steps: on host tensors (this container) the same trainer runs eagerly
    and reproduces the graph=False trainer exactly - the switch cannot change results where no graph exists."""
    from adnm_unet_b200.trainer import DataParallelTrainer
    from adnm_unet_b200.refhost import ADAMW
    g = torch.Generator().manual_seed(9)
    X, Y = torch.randn(4, 8, 6, generator=g), torch.randn(4, 8, 3, generator=g)
    finals = []
    for graph in (False, True):
        torch.manual_seed(0)
        net = Net()
        tr = DataParallelTrainer(net, _loss, clip_norm=0.05, adamw=dict(ADAMW), bucket_bytes=512, autocast_dtype=None, step_tail=torch_tail,
                                 graph=graph)
        losses = [float(tr.step(X[s], Y[s])) for s in range(4)]
        assert tr._graph is None and tr.graph_error is None
        finals.append((losses, tr.flat_p.clone()))
    assert finals[0][0] == finals[1][0] and torch.equal(finals[0][1], finals[1][1])
    # a batch of another shape after the first steps is simply another eager step
    losses = float(tr.step(X[0, :5], Y[0, :5]))
    assert losses == losses


def test_bridge_conv_layer_falls_through_on_cpu():
    """The Conv2dLayer subclass bound for the EncoderToDecoder bridges runs the reference's own forward for CPU tensors (and for any
    configuration that is not a 4-channel-group conv): same parameters, same output."""
    import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference not mounted")
    from adnm_unet_b200 import convstage
    ns = ref_loader.load_reference()
    cls = convstage.make_bridge_conv_layer(ns.ref_Conv2dLayer)
    kw = dict(in_channels=32, out_channels=32, kernel_size=(3, 1), stride=(1, 1), padding=(1, 0), bias=True, groups=8, act_func=nn.GELU)
    torch.manual_seed(1)
    a = cls(**kw)
    torch.manual_seed(1)
    b = ns.ref_Conv2dLayer(**kw)
    assert list(a.state_dict()) == list(b.state_dict()) and all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    x = torch.randn(2, 32, 6, 6)
    assert torch.equal(a(x), b(x)) and issubclass(cls, ns.ref_Conv2dLayer)
