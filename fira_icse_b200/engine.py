"""CUDA-graph training engine: the whole step (zero grads -> TransModel forward -> backward ->
[all-reduce] -> Adam) is captured once and replayed, so the ~900 launches of a step cost one
cudaGraphLaunch instead of ~900 Python/ctypes round trips (the bf16 step is launch-bound otherwise).

What makes the path capturable
  * every kernel is launched on the current stream through the C ABI, with caller-owned buffers
    (they come from the graph's private pool during capture) and no host synchronisation;
  * inputs live in STATIC device buffers (ids, shifted labels, CSR arrays with a fixed edge capacity --
    the kernels only walk rowptr ranges, so the tail of col/val is never read);
  * dropout masks are keyed by `seed + *seed_ctr`; the graph bumps the device counter on every replay,
    so replays draw fresh masks although the host-side seed is frozen into the graph;
  * TMA tensor maps are kernel parameters, rebuilt at capture time for the pooled buffers.
N > 1: the forward/backward graph, an eager NCCL all-reduce of the flat gradient bucket, then the
optimizer graph (same numerics as parallel.DataParallelStep).
"""
import torch
import torch.distributed as dist

from .graph import PackedEdges
from .parallel import FlatGradBucket

ID_KEYS = ("sou", "tar", "mark", "ast_change", "tar_label", "sub_token")


class GraphedTrainStep:
    def __init__(self, model, batch_size, optimizer_factory, edge_capacity=None, n_nodes=650, group=None,
                 lens=(210, 30, 210, 280, 30, 160)):
        self.model, self.B, self.N, self.group = model, batch_size, n_nodes, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = next(model.parameters()).device
        self.dev = dev
        self.cap = edge_capacity or batch_size * 4096
        self.ids = {k: torch.zeros((batch_size, n), dtype=torch.int64, device=dev) for k, n in zip(ID_KEYS, lens)}
        self.rowptr = torch.zeros(batch_size * n_nodes + 1, dtype=torch.int32, device=dev)
        self.col = torch.zeros(self.cap, dtype=torch.int32, device=dev)
        self.val = torch.zeros(self.cap, dtype=torch.float32, device=dev)
        self.n_global = torch.ones(1, dtype=torch.float32, device=dev)      # global token count (all ranks)
        self.seed_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        model.encoder.seed_ctr = model.decoder.seed_ctr = self.seed_ctr
        self.bucket = FlatGradBucket(model.live_parameters())
        self.optimizer = optimizer_factory(self.bucket.params)
        self.loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
        self.n_local = torch.zeros((), dtype=torch.int64, device=dev)
        self.graph_fb = self.graph_opt = None

    # ------------------------------------------------------------------ data
    def load(self, batch):
        """batch: [sou, tar, attr, mark, ast_change, edges, tar_label, sub_token] with `edges` a PackedEdges or a
        host/device (rowptr, col, val) triple.  Copies into the static buffers (async when sources are pinned)."""
        src = dict(zip(("sou", "tar", "attr", "mark", "ast_change", "edges", "tar_label", "sub_token"), batch))
        for k in ID_KEYS:
            self.ids[k].copy_(src[k], non_blocking=True)
        e = src["edges"]
        rowptr, col, val = (e.rowptr, e.col, e.val) if isinstance(e, PackedEdges) else e
        if col.numel() > self.cap:
            raise ValueError(f"batch has {col.numel()} edges, graph capacity is {self.cap}")
        self.rowptr.copy_(rowptr, non_blocking=True)
        self.col[:col.numel()].copy_(col, non_blocking=True)
        self.val[:val.numel()].copy_(val, non_blocking=True)

    def _static_batch(self):
        edges = PackedEdges(self.rowptr, self.col, self.val, self.B, self.N, True)   # fresh wrapper: no cached rowsum
        i = self.ids
        return [i["sou"], i["tar"], None, i["mark"], i["ast_change"], edges, i["tar_label"], i["sub_token"]]

    # ------------------------------------------------------------------ the step
    def _forward_backward(self):
        self.seed_ctr.add_(1)
        self.bucket.zero()
        loss_sum, n_tok = self.model(*self._static_batch(), "train")
        self.loss_sum.copy_(loss_sum.detach())
        self.n_local.copy_(n_tok)
        denom = self.n_global.squeeze(0) if self.world > 1 else n_tok.to(torch.float32)
        (loss_sum / denom).backward()

    def capture(self, warmup=3):
        """Warm up on a side stream (lazy inits, cudaFuncSetAttribute, allocator), then capture."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._count_tokens_eager()
                self._forward_backward()
                self._reduce()
                self.optimizer.step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.bucket.zero()
        self.graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fb):
            self._forward_backward()
            if self.world == 1:
                self.optimizer.step()
        if self.world > 1:
            # the captured backward always writes the same pool tensors; pack them into ONE static flat buffer
            # (eager concat + NCCL all-reduce), and let the captured optimizer read views of that buffer
            self.graph_grads = [p.grad for p in self.bucket.params]
            self.static_flat = torch.cat([g.reshape(-1) for g in self.graph_grads])
            off = 0
            for p in self.bucket.params:
                p.grad = self.static_flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                self.optimizer.step()
        return self

    def _count_tokens_eager(self):
        if self.world > 1:
            lab = self.ids["tar_label"]
            self.n_global.copy_((lab[:, 1:] != 0).sum().to(torch.float32).reshape(1))
            dist.all_reduce(self.n_global, group=self.group)

    def _reduce(self):
        if self.world > 1:
            if self.graph_fb is None:
                self.bucket.all_reduce(self.group)                   # eager warm-up iterations
            else:
                torch.cat([g.reshape(-1) for g in self.graph_grads], out=self.static_flat)
                dist.all_reduce(self.static_flat, group=self.group)

    def step(self, batch=None):
        """One training step on `batch` (or on whatever is in the static buffers).  Returns the device
        scalars (sum of the local NLL, local token count); nothing synchronises."""
        if batch is not None:
            self.load(batch)
        if self.graph_fb is None:
            self.capture()
        self._count_tokens_eager()
        self.graph_fb.replay()
        if self.world > 1:
            self._reduce()
            self.graph_opt.replay()
        return self.loss_sum, self.n_local
