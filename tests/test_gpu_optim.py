"""optim.FlatAdam (fira_adam_flat): the update of torch.optim.Adam on re-homed parameters, the bf16 mirror, zero-copy
row concatenations and in-place gradient destinations; the engine's FlatAdam paths against the reference's loss curve."""
import copy
import os

import numpy as np
import pytest
import torch

from fira_testlib import GOLDEN, seeded_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(256, 256), (256,), (256, 256), (256,), (1024, 256), (1024,), (2,), (1,), (1, 256), (333, 256), (24650,)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]


def test_flat_adam_matches_torch_adam():
    from fira_icse_b200 import optim
    from fira_icse_b200.optim import FlatAdam
    ref, mine = _params(0), _params(0)
    opt_ref = torch.optim.Adam(ref, lr=1e-3)
    opt = FlatAdam(mine, lr=1e-3, groups=[[mine[0], mine[2]], [mine[1], mine[3]]])
    # re-homing keeps values and shapes; grouped parameters are adjacent
    for a, b in zip(ref, mine):
        assert torch.equal(a, b)
    wcat = optim.cat_rows((mine[0], mine[2]))
    assert wcat.data_ptr() == mine[0].data_ptr() and wcat.shape == (512, 256)
    assert torch.equal(wcat, torch.cat((mine[0].data, mine[2].data), 0))
    assert optim.cat_rows((mine[0], mine[4])).data_ptr() != mine[0].data_ptr()       # not adjacent -> a real concatenation
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        grads = [torch.randn(p.shape, generator=g).to(DEV) * (10.0 ** (step - 3)) for p in ref]
        opt.zero_grad()
        for p, q, gr in zip(ref, mine, grads):
            p.grad = gr.clone()
            if step % 2 == 0:
                dst = optim.grad_dest((q,), tuple(q.shape))      # what the backward passes do: write in place
                assert dst is not None and optim.grad_dest((q,), tuple(q.shape)) is None     # handed out once per zero_grad
                dst.copy_(gr)
                q.grad = dst
            else:
                q.grad = gr.clone()                               # a gradient produced elsewhere: gathered by step()
        opt_ref.step()
        opt.step()
        for p, q in zip(ref, mine):
            torch.testing.assert_close(q.data, p.data, rtol=2e-6, atol=2e-7)
        # the bf16 mirror follows the parameters
        for q in mine:
            m = optim.mirror_of(q)
            assert m is not None and torch.equal(m, q.data.to(torch.bfloat16))
        assert torch.equal(optim.mirror_of(wcat), wcat.to(torch.bfloat16))
    sd = opt.state_dict()
    assert sd["step"] == 6.0
    for t, s in zip(sd["exp_avg"], opt_ref.state_dict()["state"].values()):
        torch.testing.assert_close(t, s["exp_avg"], rtol=1e-5, atol=1e-5)        # gradients up to 1e2, six decayed sums


def test_flat_adam_grad_scale():
    from fira_icse_b200.optim import FlatAdam
    ref, mine = _params(2), _params(2)
    opt_ref = torch.optim.Adam(ref, lr=1e-3)
    opt = FlatAdam(mine, lr=1e-3)
    opt.grad_scale = torch.tensor(37.0, device=DEV)
    g = torch.Generator().manual_seed(3)
    for _ in range(3):
        grads = [torch.randn(p.shape, generator=g).to(DEV) for p in ref]
        for p, q, gr in zip(ref, mine, grads):
            p.grad = gr / 37.0
            q.grad = gr.clone()
        opt_ref.step()
        opt.step()
    for p, q in zip(ref, mine):
        torch.testing.assert_close(q.data, p.data, rtol=5e-6, atol=5e-7)


def test_state_dict_roundtrip_marks_mirror_stale():
    from fira_icse_b200 import optim
    from fira_icse_b200.optim import FlatAdam
    m = copy.deepcopy(seeded_model()).to(DEV)
    keys = list(m.state_dict().keys())
    opt = FlatAdam(m.live_parameters(), lr=1e-4, groups=m.flat_groups())
    optim.attach(m, [opt])
    assert list(m.state_dict().keys()) == keys                      # the 338 reference keys, same order
    sd = {k: v.clone() + 0.5 for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    assert not opt.fresh
    optim.ensure_fresh(m)
    w = m.out_fc.weight
    assert torch.equal(optim.mirror_of(w), w.data.to(torch.bfloat16))
    # the cross-attention k/v projections of all layers are ONE [3072, 256] view
    kv = optim.cat_rows([t for c in m.decoder.cross_attention_list for t in (c.fc_k.weight, c.fc_v.weight)])
    assert kv.shape == (3072, 256) and kv.data_ptr() == m.decoder.cross_attention_list[0].fc_k.weight.data_ptr()


@pytest.mark.parametrize("mode", ["eager", "graph", "graph_split", "graph_single_opt"])
def test_training_curve_flat_adam_matches_reference(mode, monkeypatch):
    """32 Adam steps of the packed layout with optim.FlatAdam against the UNMODIFIED reference's loss curve"""
    from fira_icse_b200.engine import GraphedTrainStep
    from fira_icse_b200.optim import FlatAdam
    from fira_icse_b200.parallel import DataParallelStep
    from fira_icse_b200.packed import PackedTables, pack_from_dataset
    from test_packed import GoldenSplit, V
    gold = np.load(os.path.join(GOLDEN, "train_curve.npz"))
    m = copy.deepcopy(seeded_model()).to(DEV)
    m.eval()
    B = int(gold["batch"])
    tables = PackedTables(GoldenSplit())
    hosts = [pack_from_dataset(tables, np.arange(lo, lo + B), V) for lo in range(0, 128, B)]
    factory = lambda ps: FlatAdam(ps, lr=float(gold["lr"]), groups=m.flat_groups())      # noqa: E731
    losses = []
    if mode == "eager":
        opt = factory(m.live_parameters())
        from fira_icse_b200 import optim
        optim.attach(m, [opt])
        for _ in range(int(gold["passes"])):
            for pb in hosts:
                opt.zero_grad()
                ls, nt = m.forward_packed(pb.to(DEV), "train")
                loss = ls / nt
                loss.backward()
                assert opt.gather_grads() == 0          # every gradient was written in place by the backward passes
                opt.step()
                losses.append(loss.item())
    else:
        if mode == "graph_single_opt":
            monkeypatch.setenv("FIRA_OPT_OVERLAP", "0")
        eng = GraphedTrainStep(m, B, factory, edge_capacity=65536, split=True if mode == "graph_split" else None)
        for _ in range(int(gold["passes"])):
            for pb in hosts:
                ls, nt = eng.step(pb)
                losses.append((ls / nt).item())
    np.testing.assert_allclose(np.array(losses), gold["loss"], rtol=2e-4)
