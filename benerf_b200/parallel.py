"""Data-parallel plumbing of the training step (one process per GPU, torch.distributed).

The path shards by PIXEL: rank g renders every virtual pose of its own slice of the blur pixels and both poses of its
slice of the event pixels with a full replica of both networks, so the blur mean and the event difference stay local.
The only exchanges per step are
  * ONE all-reduce (sum) of the flat gradient buffer -- 2 x 595,844 NeRF parameters + 24 knot + 6 transform values
    (4.77 MB fp32) -- averaged over ranks (every loss is a mean over equally sized shards, loss/imgloss.py:5);
  * for E2NeRF_Real's normalised event loss (train.py:238-292) the squared norms of the predicted and measured
    event vectors, which the reference takes over the WHOLE ray batch (dim 0): a differentiable scalar all-reduce.
The reference itself is single-process (train.py:486); this module is what lets its loop run on N GPUs unchanged.
"""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(t, rank_=None, world_=None):
    """Rank's contiguous slice of a per-pixel tensor (leading dimension); shards are equal-sized (the tail is dropped
    exactly like the reference's own `sampling_rgb_rays // num_interpolated_pose`, model/nerf.py:224)."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    per = t.shape[0] // w
    return t[r * per:(r + 1) * per]


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x; dL/dx = sum over ranks of dL/dy (every rank's loss depends on every rank's x)."""

    @staticmethod
    def forward(ctx, x):
        y = x.clone()
        dist.all_reduce(y)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        dist.all_reduce(g)
        return g


def global_sum(x):
    """Differentiable sum over ranks (identity in a single process)."""
    return _AllReduceSum.apply(x) if world() > 1 else x


class FlatGrads:
    """All gradients of `params` live in ONE flat fp32 buffer (p.grad are views of it): the per-step exchange is a
    single all-reduce with no packing copies."""

    def __init__(self, params):
        self.params = list(params)
        if not all(p.requires_grad for p in self.params):
            # the flat layout is shared with the optimiser's parameter / moment buffers (train.Trainer): dropping a frozen tensor
            # would shift every later offset
            raise ValueError("FlatGrads lays out ALL given parameters; freeze an optimiser through its flag, not with requires_grad=False")
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            p._bnrf_direct_grad = True          # Graph.render's backward may add straight into p.grad (nerf._RenderFn.backward)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_sum(self):
        """Sum over ranks on the current stream (the fused optimiser tail applies the 1 / world factor itself)."""
        if world() > 1:
            dist.all_reduce(self.flat)
        return self.flat

    def all_reduce_mean(self):
        """Sum over ranks on the current stream, divided by the world size; returns the async work handle's result."""
        if world() > 1:
            dist.all_reduce(self.flat)
            self.flat.div_(world())
        return self.flat
