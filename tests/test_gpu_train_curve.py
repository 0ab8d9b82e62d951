"""End-to-end training parity on real commits: 32 Adam steps of the CUDA path (fp32 parity mode) against the loss curve
of the UNMODIFIED reference (tests/golden/train_curve.npz, written by tests/golden/make_golden_train.py from
/root/reference: same seeded weights, same batches, torch.optim.Adam lr 1e-4, dropout off).  Every step's loss within
2e-4 relative -- the forward agrees to ~1e-6, the rest is Adam's g / sqrt(v) amplifying fp32 summation-order noise on
near-zero gradient entries over 32 updates.  Also run for the per-commit PACKED layout and through the CUDA-graph engine."""
import copy
import os

import numpy as np
import pytest
import torch

from fira_testlib import GOLDEN, golden_batch, seeded_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 2e-4


@pytest.fixture(scope="module")
def gold():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return np.load(os.path.join(GOLDEN, "train_curve.npz"))


def _model():
    m = copy.deepcopy(seeded_model()).to(DEV)
    m.eval()                      # dropout off, like the golden run
    return m


def test_training_curve_matches_reference(gold):
    m = _model()
    opt = torch.optim.Adam(m.parameters(), lr=float(gold["lr"]))
    B = int(gold["batch"])
    batches = [[t.to(DEV) for t in golden_batch(lo, lo + B)] for lo in range(0, 128, B)]
    losses = []
    for _ in range(int(gold["passes"])):
        for b in batches:
            ls, nt = m(*b, "train")
            loss = ls / nt
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
    np.testing.assert_allclose(np.array(losses), gold["loss"], rtol=RTOL)


@pytest.mark.parametrize("graph", [False, True])
def test_training_curve_packed_layout_matches_reference(gold, graph):
    from fira_icse_b200.engine import GraphedTrainStep
    from fira_icse_b200.packed import PackedTables, pack_from_dataset
    from test_packed import GoldenSplit, V
    m = _model()
    B = int(gold["batch"])
    tables = PackedTables(GoldenSplit())
    hosts = [pack_from_dataset(tables, np.arange(lo, lo + B), V) for lo in range(0, 128, B)]
    losses = []
    if graph:
        eng = GraphedTrainStep(m, B, lambda ps: torch.optim.Adam(ps, lr=float(gold["lr"]), fused=True, capturable=True),
                               edge_capacity=65536)
        for _ in range(int(gold["passes"])):
            for pb in hosts:
                ls, nt = eng.step(pb)
                losses.append((ls / nt).item())
    else:
        opt = torch.optim.Adam(m.parameters(), lr=float(gold["lr"]))
        for _ in range(int(gold["passes"])):
            for pb in hosts:
                ls, nt = m.forward_packed(pb.to(DEV), "train")
                loss = ls / nt
                opt.zero_grad()
                loss.backward()
                opt.step()
                losses.append(loss.item())
    np.testing.assert_allclose(np.array(losses), gold["loss"], rtol=RTOL)
