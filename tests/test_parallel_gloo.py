"""world_size-2 gloo test of the data-parallel step logic (host side of parallel.py): the all-reduced
gradient and the loss must equal a single-process step on the concatenated batch, with the reference's
global token-weighted normalisation sum(loss)/sum(tokens) (run_model.py:105) -- not a mean of means."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class Toy(nn.Module):
    """Stands in for TransModel on CPU: returns (loss_sum, n_tokens) like Model.py:84."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = nn.Linear(8, 8)
        self.b = nn.Linear(8, 1)
        self.dead = nn.Linear(8, 8)          # never used: must stay out of the bucket

    def live_parameters(self):
        return list(self.a.parameters()) + list(self.b.parameters())

    def forward(self, x, w, stage="train"):
        per_tok = self.b(torch.tanh(self.a(x))).squeeze(-1) ** 2
        return (per_tok * w).sum(), (w != 0).sum()


def _data():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(10, 8, generator=g)
    w = torch.tensor([1, 1, 1, 1, 1, 1, 1, 0, 0, 0.0])          # ranks see 5 and 2 tokens: mean-of-means differs
    return x, w


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fira_icse_b200.parallel import DataParallelStep
    x, w = _data()
    lo, hi = rank * 5, rank * 5 + 5
    dp = DataParallelStep(Toy(), lambda ps: torch.optim.SGD(ps, lr=0.1))
    loss, n = dp.step([x[lo:hi], w[lo:hi]])
    # plain python lists: tensors sent through a spawn Queue die with the sending process
    q.put((rank, loss.item(), n.item(), torch.cat([p.grad.reshape(-1) for p in dp.bucket.params]).tolist(), [p.detach().flatten().tolist() for p in dp.bucket.params],
           dp.model.dead.weight.grad is None))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.timeout(300)
def test_two_rank_step_equals_single_process_step():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single process, whole batch
    x, w = _data()
    m = Toy()
    loss_sum, n = m(x, w)
    (loss_sum / n).backward()
    ref_flat = torch.cat([p.grad.flatten() for p in m.live_parameters()])
    opt = torch.optim.SGD(m.live_parameters(), lr=0.1)
    opt.step()
    for rank, loss, ntok, flat, params, dead_none in res:
        assert ntok == 7
        assert abs(loss - (loss_sum / n).item()) < 1e-6
        assert torch.allclose(torch.tensor(flat), ref_flat, atol=1e-6)
        assert dead_none
        for p, q_ in zip(params, m.live_parameters()):
            assert torch.allclose(torch.tensor(p), q_.detach().flatten(), atol=1e-6)
    assert res[0][3] == res[1][3]                     # replicas stay bit-identical
