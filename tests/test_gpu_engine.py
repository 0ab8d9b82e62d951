"""CUDA-graph training engine: replaying the captured step must equal the eager step, and replays must
draw fresh dropout masks (device-side seed counter)."""
import copy

import numpy as np
import pytest
import torch

from fira_testlib import golden_batch, seeded_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _packed(lo, hi):
    from fira_icse_b200 import PackedEdges
    b = golden_batch(lo, hi, dense_edge=False)
    b[5] = PackedEdges.from_coo_lists(b[5], 650, DEV)
    return [x.to(DEV) if torch.is_tensor(x) else x for x in b]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graph_replay_equals_eager_training(precision):
    from fira_icse_b200.engine import GraphedTrainStep
    from fira_icse_b200.parallel import DataParallelStep
    base = copy.deepcopy(seeded_model()).to(DEV).set_precision(precision)
    base.eval()                                   # dropout off -> deterministic comparison
    m_eager, m_graph = copy.deepcopy(base), copy.deepcopy(base)
    batches = [_packed(i * 8, i * 8 + 8) for i in range(4)]
    # plain SGD for the equality check: Adam's g/sqrt(v) turns fp32 summation-order noise on near-zero
    # gradient entries into full +-lr steps (the Adam graph path is covered by the test below and by bench.py)
    dp = DataParallelStep(m_eager, lambda ps: torch.optim.SGD(ps, lr=2e-3))
    eng = GraphedTrainStep(m_graph, 8, lambda ps: torch.optim.SGD(ps, lr=2e-3))
    # capture() runs ONE real (eager) step on whatever is loaded before capturing: replicate it eagerly
    eng.load(batches[0])
    eng.capture()
    dp.step(batches[0])
    tol = 5e-3 if precision == "bf16" else 2e-4
    for b in batches:
        loss_e, _ = dp.step(b)
        ls, n = eng.step(b)
        loss_g = (ls / n).item()
        assert abs(loss_g - loss_e.item()) <= tol * abs(loss_e.item()), (loss_g, loss_e.item())
    for (k, p), (_, q) in zip(m_eager.named_parameters(), m_graph.named_parameters()):
        if p.grad is None:
            continue
        assert torch.allclose(p, q, rtol=0, atol=5e-3 if precision == "bf16" else 5e-4), k


def test_graph_replays_draw_fresh_dropout_masks():
    from fira_icse_b200.engine import GraphedTrainStep
    m = copy.deepcopy(seeded_model()).to(DEV)
    m.train()
    eng = GraphedTrainStep(m, 8, lambda ps: torch.optim.Adam(ps, lr=0.0, fused=True, capturable=True))
    b = _packed(0, 8)
    eng.load(b)
    eng.capture()
    c0 = int(eng.seed_ctr.item())
    losses = []
    for _ in range(4):
        ls, n = eng.step(b)
        losses.append((ls / n).item())
    assert int(eng.seed_ctr.item()) == c0 + 4
    assert len({round(x, 6) for x in losses}) == 4, losses      # lr = 0: only the masks differ
    assert np.std(losses) < 0.05 * np.mean(losses)


def test_multi_shape_graphs_share_one_pool_and_match_eager():
    """Trimmed batches of different (batch, n_code, n_sub, n_ast) shapes: one graph per shape captured into a
    shared memory pool, replayed in arbitrary order, must train exactly like the eager step."""
    from fira_icse_b200 import PackedEdges
    from fira_icse_b200.data import trim_batch_host
    from fira_icse_b200.engine import GraphedTrainStep
    from fira_icse_b200.parallel import DataParallelStep
    base = copy.deepcopy(seeded_model()).to(DEV)
    base.eval()
    m_eager, m_graph = copy.deepcopy(base), copy.deepcopy(base)

    def host(lo, hi):
        b = golden_batch(lo, hi, dense_edge=False)
        full = [b[0], b[1], None, b[3], b[4], PackedEdges.pack_host(b[5], 650), b[6], b[7]]
        return trim_batch_host(full, base.vocab_size)

    def on_device(h):
        n, nodes = h[0].shape[0], h[0].shape[1] + h[7].shape[1] + h[4].shape[1]
        d = [x.to(DEV) if torch.is_tensor(x) else x for x in h]
        d[5] = PackedEdges.from_host(*h[5], n, nodes, DEV)
        return d
    hosts = [host(0, 8), host(8, 16), host(40, 48), host(100, 105)]          # the last one is a short batch
    shapes = {(h[0].shape[0], h[0].shape[1], h[7].shape[1], h[4].shape[1]) for h in hosts}
    assert len(shapes) == 4
    dp = DataParallelStep(m_eager, lambda ps: torch.optim.SGD(ps, lr=2e-3))
    eng = GraphedTrainStep(m_graph, 8, lambda ps: torch.optim.SGD(ps, lr=2e-3))
    for k in [0, 1, 2, 3, 0, 2, 1, 3, 3, 0, 1]:
        loss_e, _ = dp.step(on_device(hosts[k]))
        ls, n = eng.step(hosts[k])
        loss_g = (ls / n).item()
        assert abs(loss_g - loss_e.item()) <= 2e-4 * abs(loss_e.item()), (k, loss_g, loss_e.item())
    assert len(eng.captured) == 4 and all(c.graph is not None for c in eng.captured.values())
    assert eng.pool is not None
    for (k, p), (_, q) in zip(m_eager.named_parameters(), m_graph.named_parameters()):
        if p.grad is None:
            continue
        assert torch.allclose(p, q, rtol=0, atol=5e-4), k


def test_split_graphs_equal_the_single_graph_step():
    """engine.GraphedTrainStep(split=True): graph A (forward + head/decoder backward) + graph B (encoder backward) +
    optimizer graph -- the N > 1 path that overlaps the gradient all-reduce with graph B -- on one GPU (no NCCL) trains
    exactly like the eager step"""
    from fira_icse_b200.engine import GraphedTrainStep
    from fira_icse_b200.parallel import DataParallelStep
    base = copy.deepcopy(seeded_model()).to(DEV)
    base.eval()
    m_eager, m_graph = copy.deepcopy(base), copy.deepcopy(base)
    batches = [_packed(i * 8, i * 8 + 8) for i in range(3)]
    dp = DataParallelStep(m_eager, lambda ps: torch.optim.SGD(ps, lr=2e-3))
    eng = GraphedTrainStep(m_graph, 8, lambda ps: torch.optim.SGD(ps, lr=2e-3), split=True)
    eng.load(batches[0])
    eng.capture()
    dp.step(batches[0])
    for b in batches + batches:
        loss_e, _ = dp.step(b)
        ls, n = eng.step(b)
        assert abs((ls / n).item() - loss_e.item()) <= 2e-4 * abs(loss_e.item())
    assert eng.params_a and eng.params_b and len(eng.params_a) + len(eng.params_b) == len(eng.bucket.params)
    names = {id(p): k for k, p in m_graph.named_parameters()}
    assert all(names[id(p)].startswith("encoder.") for p in eng.params_b)
    assert not any(names[id(p)].startswith("encoder.") for p in eng.params_a)
    for (k, p), (_, q) in zip(m_eager.named_parameters(), m_graph.named_parameters()):
        if p.grad is None:
            continue
        assert torch.allclose(p, q, rtol=0, atol=5e-4), k


@pytest.mark.parametrize("split", [False, True])
def test_packed_batches_through_the_graph_engine(split):
    """per-commit packed batches (packed.PackedBatch) replayed as CUDA graphs == eager TransModel.forward_packed"""
    import sys
    from fira_icse_b200.engine import GraphedTrainStep
    from fira_icse_b200.packed import PackedTables, pack_from_dataset
    from test_packed import GoldenSplit, V
    tables = PackedTables(GoldenSplit())
    hosts = [pack_from_dataset(tables, np.arange(lo, lo + 8), V) for lo in (0, 8, 40)]
    base = copy.deepcopy(seeded_model()).to(DEV)
    base.eval()
    m_eager, m_graph = copy.deepcopy(base), copy.deepcopy(base)
    opt = torch.optim.SGD(m_eager.live_parameters(), lr=2e-3)
    eng = GraphedTrainStep(m_graph, 8, lambda ps: torch.optim.SGD(ps, lr=2e-3), edge_capacity=32768, split=split)

    def eager(pb):
        opt.zero_grad(set_to_none=True)
        ls, nt = m_eager.forward_packed(pb.to(DEV), "train")
        loss = ls / nt
        loss.backward()
        opt.step()
        return loss.item()
    for k in [0, 1, 2, 0, 2, 1]:
        le = eager(hosts[k])
        ls, n = eng.step(hosts[k])
        assert abs((ls / n).item() - le) <= 2e-4 * abs(le), (k, (ls / n).item(), le)
    for (k, p), (_, q) in zip(m_eager.named_parameters(), m_graph.named_parameters()):
        if p.grad is None:
            continue
        assert torch.allclose(p, q, rtol=0, atol=5e-4), k
