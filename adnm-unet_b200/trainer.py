"""Batch-sharded data-parallel training step for ADNM-UNet: one process per GPU, NCCL all-reduce of the gradients only.

Replaces the reference's `nn.DataParallel` wrap (train.py:99-102) and the tail of its step (train.py:140-145).  Every op
of the network is per-sample and the loss is sum/numel over the local shard (models/loss.py:64-65), so
equal shards + gradient averaging reproduce the single-GPU big-batch step exactly.

What is B200-specific here:
  * the 669 parameter tensors that actually receive a gradient (of 992; the other 307 + the 16 frozen Haar filters are
    discovered on the first step and left alone - AdamW skips `grad is None`, so no decay is applied to them either)
    live in ONE flat fp32 buffer, in the order their gradients become ready during backward; so do their gradients and
    the two Adam moments;
  * the gradient buffer is cut into buckets; a post-accumulate hook counts a bucket's tensors down and enqueues its
    NCCL all-reduce (sum over NVLink / NVSwitch) as soon as the last one is written - the reduction of the refiner's
    gradients overlaps the backward of the decoder and encoder;
  * global grad-norm, clip (train.py:140), AdamW (train_untils.py:35-42) and zero_grad are two bandwidth-bound passes
    over the flat buffers in the sm_100a library (include/adnb200.h: adn_sumsq_f32, adn_adamw_flat), with the 1/world
    average folded in; no `.item()` host synchronisation per step (the reference has two, train.py:141,146);
  * `graph=True`: forward + loss + backward are captured ONCE into a CUDA graph (every library entry point enqueues on the
    caller's stream without allocation or synchronisation, and the hosted reference modules are plain tensor ops) and
    replayed per step on static input buffers - the ~1000 library launches + ~2000 tensor-op launches of a step then cost
    one `cudaGraphLaunch` instead of ~20 ms of host time, which is what bounds the eager step at B = 32 per GPU.  The
    all-reduce (bucketed, NCCL) and the two clip / AdamW kernels stay outside the graph: the step counter and the learning
    rate are host scalars that change every step.
"""
import math

import torch
import torch.distributed as dist

from adnm_unet_b200 import _lib
from adnm_unet_b200.refhost import ADAMW, CLIP_NORM


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous equal shards; the reference loss is sum/numel (models/loss.py:64-65), so shards must be equal-sized."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def reference_lr(epoch: int, base_lr=1e-3, warmup_epochs=3, t_max=50, eta_min=5e-7):
    """train_untils.py:44-46: LinearLR(start 0.01, 3 epochs) then CosineAnnealingLR(T_max 50, eta_min 5e-7); `epoch` counts
    completed lr_scheduler.step() calls (train.py:187)."""
    if epoch < warmup_epochs:
        return base_lr * (0.01 + (1.0 - 0.01) * epoch / warmup_epochs)
    e = epoch - warmup_epochs
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * e / t_max)) / 2


def cuda_step_tail(p, g, m, v, ws, step, lr, grad_scale, clip_norm, hp):
    """clip_grad_norm_ + AdamW.step + zero_grad on the flat buffers, in the sm_100a library.  No CPU implementation."""
    _lib.require_cuda(p, "flat parameter buffer")
    lib = _lib.load()
    n = p.numel()
    st = _lib.stream_ptr(p.device)
    with _lib.on_device(p.device):
        _lib.check(lib.adn_sumsq_f32(_lib.ptr(g), n, _lib.ptr(ws["partial"]), _lib.ptr(ws["sumsq"]), st), "adn_sumsq_f32")
        _lib.check(lib.adn_adamw_flat(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), n, _lib.ptr(ws["sumsq"]),
                                      _lib.ptr(ws["norm"]), lr, hp["beta1"], hp["beta2"], hp["eps"], hp["weight_decay"],
                                      step, grad_scale, clip_norm if clip_norm else 0.0, st), "adn_adamw_flat")


class DataParallelTrainer:
    """model: the (drop-in hosting) network; loss_fn(outputs, targets) -> scalar.

    step(imgs, targets, lr=None) runs forward, loss, backward (bucketed all-reduce overlapped), clip, AdamW, zero_grad and
    returns the local loss as a device tensor.  The very first step also discovers which parameters are live and builds
    the flat buffers (so it is slower and its all-reduce is not overlapped)."""

    def __init__(self, model, loss_fn, group=None, clip_norm=CLIP_NORM, adamw=None, bucket_bytes=32 << 20,
                 autocast_dtype=torch.bfloat16, step_tail=None, graph=False):
        self.model, self.loss_fn, self.group = model, loss_fn, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.clip_norm, self.hp = clip_norm, dict(ADAMW if adamw is None else adamw)
        self.bucket_bytes, self.autocast_dtype = int(bucket_bytes), autocast_dtype
        self.step_tail = cuda_step_tail if step_tail is None else step_tail
        self.steps_done = 0
        self.live = None            # parameters with gradients, in gradient-ready order
        self.flat_p = self.flat_g = self.flat_m = self.flat_v = None
        self.buckets = []           # (start, end) element ranges of flat_g
        self._pending, self._works, self._ready_order = [], [], []
        self._hooks = []
        self.graph = bool(graph)      # capture forward + loss + backward into a CUDA graph after the discovery step
        self._graph, self._static, self._eager_after_discovery, self._stream = None, None, 0, None
        self.graph_error = None       # why the capture was abandoned (the trainer then stays eager), for the bench line

    # ------------------------------------------------------------------ forward / backward
    def _forward_loss(self, imgs, targets):
        dev = imgs.device.type
        if self.autocast_dtype is not None and dev == "cuda":
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                out = self.model(imgs)
            return self.loss_fn(out.float(), targets)
        return self.loss_fn(self.model(imgs), targets)

    # ------------------------------------------------------------------ discovery + flattening (first step)
    def _discover_and_flatten(self, imgs, targets):
        params = [p for p in self.model.parameters() if p.requires_grad]
        order = []
        hooks = [p.register_post_accumulate_grad_hook(lambda p, order=order: order.append(p)) for p in params]
        for p in params:
            p.grad = None
        loss = self._forward_loss(imgs, targets)
        loss.backward()
        for h in hooks:
            h.remove()
        live = [p for p in order if p.grad is not None]
        # every rank must agree on the live set (it is structural: same model, any data) - checked, not assumed
        index = {id(p): i for i, p in enumerate(params)}
        mask = torch.zeros(len(params), device=imgs.device)
        mask[[index[id(p)] for p in live]] = 1.0
        if self.world > 1:
            total = mask.clone()
            dist.all_reduce(total, group=self.group)
            if not torch.equal(total, mask * self.world):
                raise RuntimeError("data-parallel ranks disagree on which parameters receive gradients")
        if any(p.dtype != torch.float32 for p in live):
            raise RuntimeError("the flat trainer keeps fp32 master weights; found a non-fp32 trainable parameter")
        self.live = live
        sizes = [(p.numel() + 3) // 4 * 4 for p in live]          # 16-byte aligned slots
        n = sum(sizes)
        dev = imgs.device
        self.flat_p, self.flat_g = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        self.flat_m, self.flat_v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        self.ws = {"partial": torch.zeros(4096, device=dev), "sumsq": torch.zeros(1, device=dev), "norm": torch.zeros(1, device=dev)}
        self.offsets, off = [], 0
        with torch.no_grad():
            for p, sz in zip(live, sizes):
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                self.flat_g[off:off + k].copy_(p.grad.reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                p.grad = self.flat_g[off:off + k].view(p.shape)
                self.offsets.append(off)
                off += sz
        # buckets: consecutive tensors (in ready order) up to bucket_bytes
        self.buckets, self.bucket_of, start, count = [], [], 0, 0
        for i, sz in enumerate(sizes):
            self.bucket_of.append(len(self.buckets))
            count += 1
            end = self.offsets[i] + sz
            if (end - start) * 4 >= self.bucket_bytes or i == len(sizes) - 1:
                self.buckets.append((start, end, count))
                start, count = end, 0
        self._pending = [c for _, _, c in self.buckets]
        for i, p in enumerate(live):
            self._hooks.append(p.register_post_accumulate_grad_hook(lambda p, b=self.bucket_of[i]: self._grad_ready(b)))
        self._overlap = False
        return loss

    def _grad_ready(self, b):
        if not self._overlap:
            return
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.world > 1:
            s, e, _ = self.buckets[b]
            self._works.append(dist.all_reduce(self.flat_g[s:e], group=self.group, async_op=True))

    # ------------------------------------------------------------------ CUDA-graph capture of forward + loss + backward
    def _capture(self, imgs, targets):
        """Static input buffers, then the capture on the trainer's own stream - the stream every earlier step of this trainer
        ran on, so the parameters' AccumulateGrad nodes (created on first use, kept alive by the gradient hooks) belong to the
        capturing stream; a node that lives on the legacy default stream would invalidate the capture.  The gradients
        accumulate in place into the views of `flat_g` (zeroed by the optimizer kernel at the end of every step), inside the
        graph as outside; the eager steps before this call have warmed every lazily initialised handle."""
        st_imgs, st_tgt = imgs.detach().clone(), targets.detach().clone()
        graph = torch.cuda.CUDAGraph()
        # thread_local: CUDA calls of OTHER threads (NCCL watchdog, a data loader pinning memory) must not invalidate the capture
        with torch.cuda.graph(graph, stream=self._stream, capture_error_mode="thread_local"):
            loss = self._forward_loss(st_imgs, st_tgt)
            loss.backward()
        self._graph, self._static = graph, (st_imgs, st_tgt, loss.detach())

    def _graph_step(self, imgs, targets):
        if self._graph is None:
            try:
                self._capture(imgs, targets)
            except Exception as ex:      # an op of the hosted network that cannot be captured: stay eager, say so
                self.graph, self.graph_error = False, f"{type(ex).__name__}: {ex}"[:300]
                torch.cuda.synchronize()
                self.flat_g.zero_()
                return None
        st_imgs, st_tgt, st_loss = self._static
        st_imgs.copy_(imgs, non_blocking=True)
        st_tgt.copy_(targets, non_blocking=True)
        self._graph.replay()
        if self.world > 1:
            works = [dist.all_reduce(self.flat_g[s:e], group=self.group, async_op=True) for s, e, _ in self.buckets]
            for w in works:
                w.wait()
        return st_loss

    # ------------------------------------------------------------------ one training step (train.py:132-146)
    def step(self, imgs, targets, lr=None):
        if self.graph and imgs.is_cuda:
            # graph mode: every step of this trainer runs on ONE non-default stream (see _capture)
            if self._stream is None:
                self._stream = torch.cuda.Stream(imgs.device)
            cur = torch.cuda.current_stream(imgs.device)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                loss = self._step(imgs, targets, lr)
            cur.wait_stream(self._stream)
            return loss
        return self._step(imgs, targets, lr)

    def _step(self, imgs, targets, lr=None):
        self.model.train()
        loss = None
        if self.live is None:
            loss = self._discover_and_flatten(imgs, targets)
            if self.world > 1:
                dist.all_reduce(self.flat_g, group=self.group)
        elif (self.graph and imgs.is_cuda and self._eager_after_discovery >= 1
              and (self._static is None or (imgs.shape == self._static[0].shape and targets.shape == self._static[1].shape))):
            loss = self._graph_step(imgs, targets)      # a batch of another shape (the last one of an epoch) runs eagerly below
        if loss is None:
            self._eager_after_discovery += 1
            self._pending = [c for _, _, c in self.buckets]
            self._works, self._overlap = [], True
            loss = self._forward_loss(imgs, targets)
            loss.backward()
            self._overlap = False
            if self.world > 1 and any(self._pending):
                raise RuntimeError("a live parameter received no gradient this step: the live set changed after discovery")
            for w in self._works:
                w.wait()            # stream-level wait (NCCL), not a host block
        self.steps_done += 1
        self.step_tail(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.ws, self.steps_done,
                       self.hp["lr"] if lr is None else lr, 1.0 / self.world, self.clip_norm, self.hp)
        return loss.detach()

    def grad_norm(self):
        """Unclipped global gradient norm of the last step (device tensor; what train.py:141 logs)."""
        return self.ws["norm"]

    def n_live_elements(self):
        return sum(p.numel() for p in self.live) if self.live is not None else 0
