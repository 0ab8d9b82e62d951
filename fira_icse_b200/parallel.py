"""Data parallelism: one process per GPU, graphs sharded by commit, NCCL all-reduce of gradients only.

The reference wraps the model in nn.DataParallel (run_model.py:392-394): one process, per-step
parameter broadcast + gradient reduce to GPU 0, and the loss is sum(loss) / sum(tokens) over the
gathered replicas (run_model.py:105).  Here every rank keeps its own replica and optimizer state;
per step there is ONE all-reduce over a flat gradient bucket (one concatenation packs the gradients,
afterwards the parameters' .grad tensors are views into the reduced buffer) plus an 8-byte all-reduce of the token count so that the
loss is the same global token-weighted mean as upstream.  The 74 tensors that never receive
gradients (encoder.lstm, encoder.combination_list1, gate_fc) are left out of the bucket.
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """Gradients of `params` as ONE flat fp32 buffer for the all-reduce.

    Gradients are not pre-bound to the buffer: with `.grad = None` before backward autograd adopts
    the tensors our Functions return (no zero-fill, no `grad += new` kernel per parameter -- 258 add
    launches per step in the round-1 profile).  flatten() packs them with one concatenation;
    after the all-reduce every `.grad` is re-pointed at its slice of the reduced buffer."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None

    def zero(self):
        for p in self.params:
            p.grad = None

    def flatten(self):
        self.flat = torch.cat([p.grad.reshape(-1) for p in self.params])
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        return self.flat

    def all_reduce(self, group=None, async_op=False):
        return dist.all_reduce(self.flatten(), op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class DataParallelStep:
    """loss = sum_ranks(loss_sum) / sum_ranks(tokens); grads all-reduced; Adam on every rank.

    step(batch) -> (global mean loss as a 0-dim tensor, global token count)."""

    def __init__(self, model, optimizer_factory, group=None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        live = model.live_parameters() if hasattr(model, "live_parameters") else list(model.parameters())
        self.bucket = FlatGradBucket(live)
        self.optimizer = optimizer_factory(self.bucket.params)
        from .optim import FlatAdam, attach
        if isinstance(self.optimizer, FlatAdam):
            attach(model, [self.optimizer])

    def step(self, batch, stage="train"):
        self.bucket.zero()
        self.optimizer.zero_grad(set_to_none=True)       # optim.FlatAdam: also zero-fills its flat gradient buffer
        loss_sum, n_tok = self.model(*batch, stage)
        n_global = n_tok.to(torch.float32).reshape(1).clone()
        loss_global = loss_sum.detach().reshape(1).clone()
        if self.world > 1:
            dist.all_reduce(n_global, group=self.group)
        (loss_sum / n_global.squeeze(0)).backward()
        if self.world > 1:
            self.bucket.all_reduce(self.group)
            dist.all_reduce(loss_global, group=self.group)
        self.optimizer.step()
        return (loss_global / n_global).squeeze(0), n_global.squeeze(0)


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` (inference: outputs concatenate in index order)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
