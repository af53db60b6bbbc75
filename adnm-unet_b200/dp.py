"""Batch-sharded data parallelism for the drop-in modules: one process per GPU, the only exchange is one all-reduce of the
parameter gradients per step (every op of the mixer / WTConv2d is per-sample, so equal shards +
gradient averaging reproduce the single-GPU big-batch gradient exactly).  Replaces the reference's `nn.DataParallel`
(train.py:99-102).  Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous equal shards; the reference loss is sum/numel (models/loss.py:64-65), so shards must be equal-sized."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


class GradAllReducer:
    """Flat-bucket all-reduce(sum) / world of the gradients of a FIXED parameter list.

    Parameters whose grad is None on this rank (the reference leaves 307 tensors without gradient, SURVEY.md note 7)
    contribute zeros and keep grad None afterwards only if they are None on every rank (checked once, at construction
    time of the mask, by all-reducing a presence bitmap) - so no rank ever blocks on a missing bucket."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.sizes = [p.numel() for p in self.params]
        dev = self.params[0].device if self.params else torch.device("cpu")
        # one bucket dtype: the widest parameter dtype (fp32 masters in every use of this class); narrower parameters
        # get a converted copy back as their .grad
        dt = torch.float32
        if self.params and all(p.dtype == self.params[0].dtype for p in self.params):
            dt = self.params[0].dtype
        self.flat = torch.zeros(sum(self.sizes), dtype=dt, device=dev)
        self.present = None

    def _presence(self):
        mask = torch.tensor([0.0 if p.grad is None else 1.0 for p in self.params], device=self.flat.device)
        if self.world > 1:
            dist.all_reduce(mask, group=self.group)
        self.present = (mask > 0).tolist()

    @torch.no_grad()
    def __call__(self):
        """Gather into the flat bucket (slice-wise copies; a gradient that already IS its slice - kept by the caller across
        steps, e.g. optimizer.zero_grad(set_to_none=False) or gradient accumulation - is left in place instead of being
        copied onto itself), one all-reduce, one scale; afterwards every p.grad is a VIEW into the flat bucket."""
        if self.present is None:
            self._presence()
        live = [(p, n) for p, n, pres in zip(self.params, self.sizes, self.present) if pres]
        total = sum(n for _, n in live)
        used = self.flat[:total]
        off = 0
        for p, n in live:
            dst = used[off:off + n]
            if p.grad is None:
                dst.zero_()
            elif p.grad.data_ptr() != dst.data_ptr() or p.grad.dtype != dst.dtype:
                dst.copy_(p.grad.reshape(-1))
            off += n
        if self.world > 1:
            dist.all_reduce(used, group=self.group)
            used.mul_(1.0 / self.world)
        off = 0
        for p, n in live:
            v = used[off:off + n].view_as(p)
            p.grad = v if p.dtype == used.dtype else v.to(p.dtype)
            off += n
        return used

    def grad_norm(self):
        """Global L2 norm of the reduced gradients (what train.py:140 clip_grad_norm_ needs), no extra pass over params."""
        n = sum(s for s, pres in zip(self.sizes, self.present) if pres)
        return self.flat[:n].norm()
