"""Data parallelism: one process per GPU, graphs sharded by commit, NCCL all-reduce of gradients only.

The reference wraps the model in nn.DataParallel (run_model.py:392-394): one process, per-step
parameter broadcast + gradient reduce to GPU 0, and the loss is sum(loss) / sum(tokens) over the
gathered replicas (run_model.py:105).  Here every rank keeps its own replica and optimizer state;
per step there is ONE all-reduce over a flat gradient bucket (the parameters' .grad tensors are views
into it, so there is no pack/unpack copy) plus an 8-byte all-reduce of the token count so that the
loss is the same global token-weighted mean as upstream.  The 74 tensors that never receive
gradients (encoder.lstm, encoder.combination_list1, gate_fc) are left out of the bucket.
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """Flat fp32 buffer holding the gradients of `params`; each p.grad is a view into it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None, async_op=False):
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


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

    def step(self, batch, stage="train"):
        self.bucket.zero()
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
