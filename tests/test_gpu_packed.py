"""Per-commit packed batches on the GPU: TransModel.forward_packed (node rows = real nodes only) against
TransModel.forward on the reference's padded batch -- same loss, token count, gradients and argmax ids (fp32 parity
mode: fp32 round-off; bf16 mode: the two layouts round identically row by row, only reduction orders differ)."""
import copy

import numpy as np
import pytest
import torch

from fira_testlib import golden_batch, seeded_model
from test_packed import GoldenSplit, V

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def model():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    m = copy.deepcopy(seeded_model()).to(DEV)
    m.eval()
    return m


def _packed(index):
    from fira_icse_b200.packed import PackedTables, pack_from_dataset
    return pack_from_dataset(PackedTables(GoldenSplit()), np.asarray(index), V).to(DEV)


def _padded(index):
    parts = [golden_batch(i, i + 1) for i in index]
    return [torch.cat([p[k] for p in parts], 0).to(DEV) for k in range(8)]


def _run(model, fn):
    model.zero_grad(set_to_none=True)
    ls, nt = fn()
    (ls / nt).backward()
    return ls.item(), int(nt), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("index", [list(range(0, 10)), [5], [100, 3, 77, 127, 64, 9]])
def test_packed_forward_backward_equals_padded(model, index):
    pb = _packed(index)
    pad = _padded(index)
    l_pad, n_pad, g_pad = _run(model, lambda: model(*pad, "train"))
    l_pk, n_pk, g_pk = _run(model, lambda: model.forward_packed(pb, "train"))
    assert n_pad == n_pk
    assert abs(l_pad - l_pk) <= 5e-6 * abs(l_pad), (l_pad, l_pk)
    assert sorted(g_pad) == sorted(g_pk)
    for k in g_pad:
        scale = g_pad[k].abs().max().item()
        if scale < 1e-6:
            continue
        assert (g_pad[k] - g_pk[k]).abs().max().item() <= 2e-4 * scale + 1e-9, k
    # argmax ids: vocabulary ids identical, copy ids renumbered from the padded positions to the commit's memory rows
    with torch.no_grad():
        ids_pad = model(*pad, "dev").cpu().numpy()
        ids_pk = model.forward_packed(pb, "dev").cpu().numpy()
    ranges = pb.ranges.cpu().numpy()
    for b in range(len(index)):
        uc = ranges[b][1]
        exp = ids_pad[b].copy()
        sub = exp >= V + 210
        exp[sub] = V + uc + (exp[sub] - V - 210)
        assert np.array_equal(ids_pk[b], exp), b


def test_packed_bf16_mode_matches_padded_bf16_mode(model):
    m = copy.deepcopy(model).set_precision("bf16")
    index = list(range(16, 32))
    pb, pad = _packed(index), _padded(index)
    l_pad, n_pad, g_pad = _run(m, lambda: m(*pad, "train"))
    l_pk, n_pk, g_pk = _run(m, lambda: m.forward_packed(pb, "train"))
    assert n_pad == n_pk and abs(l_pad - l_pk) <= 2e-3 * abs(l_pad), (l_pad, l_pk)
    for k in g_pad:
        a, b = g_pad[k].double().flatten(), g_pk[k].double().flatten()
        # fc_k.bias / LinearRes.bias vanish in exact arithmetic (softmax shift invariance): round-off noise
        if a.norm().item() < 1e-6 or k.endswith("fc_k.bias") or k.endswith("LinearRes.bias"):
            continue
        cos = float((a @ b) / (a.norm() * b.norm()))
        assert cos > 0.995, (k, cos)


def test_packed_rows_with_bucket_padding_are_inert(model):
    """a larger bucket (more empty rows at the end of every segment) does not change anything"""
    from fira_icse_b200.packed import PackedTables, pack_from_dataset
    index = np.arange(40, 46)
    t = PackedTables(GoldenSplit())
    a = pack_from_dataset(t, index, V).to(DEV)
    b = pack_from_dataset(t, index, V, pad_dims=(a.Rc + 1024, a.Rs + 512, a.Ra + 512, a.S + 64)).to(DEV)
    la, na, ga = _run(model, lambda: model.forward_packed(a, "train"))
    lb, nb, gb = _run(model, lambda: model.forward_packed(b, "train"))
    assert na == nb and abs(la - lb) <= 2e-6 * abs(la)
    for k in ga:
        scale = ga[k].abs().max().item()
        if scale >= 1e-6:
            assert (ga[k] - gb[k]).abs().max().item() <= 1e-4 * scale + 1e-9, k
