"""CUDA-graph replay of a mixer's forward and backward for ordinary (eager) training loops.

One fwd+bwd of the mixer is ~15 kernel / memset launches issued through Python and autograd: 0.37 ms of host time per
step at the benchmark shape, more than the 0.28 ms the GPU needs.  The C ABI never allocates or synchronises, so both
passes are capturable; `graphed_mixer` wraps `torch.cuda.make_graphed_callables` (which captures the forward and the
backward as two graphs and hooks them into autograd) for a fixed input shape."""
import torch
import torch.nn as nn


class _BoundMixer(nn.Module):
    """forward(u) of a Mamba2 with its token grid fixed (graphed callables take tensors only)."""

    def __init__(self, mixer, H, W):
        super().__init__()
        self.mixer, self.H, self.W = mixer, int(H), int(W)

    def forward(self, u):
        return self.mixer(u, self.H, self.W)


def graphed_mixer(mixer, sample_u, H, W, num_warmup_iters=3):
    """Return `f(u) -> out` that replays captured CUDA graphs of `mixer(u, H, W)` and of its backward.

    `sample_u` fixes shape, dtype and device (its values are irrelevant); every later `u` must match it.  Gradients reach
    `u` and the mixer's parameters through autograd as usual (`scale`, `shift`, `alpha2` keep `grad = None`, as in the
    reference).  Parameters are read at replay time, so optimizer updates in place are seen."""
    if not sample_u.is_cuda:
        raise RuntimeError("graphed_mixer needs a CUDA tensor (no CPU fallback)")
    bound = _BoundMixer(mixer, H, W)
    sample = sample_u.detach().clone().requires_grad_(True)
    return torch.cuda.make_graphed_callables(bound, (sample,), num_warmup_iters=num_warmup_iters, allow_unused_input=True)
