"""Product graph/batch builder (fira_icse_b200.data) against the reference's own process_data output
(tests/golden/batch_first128.npz) and against the oracle restatement; BLEU restatement sanity."""
import json
import os

import numpy as np
import pytest
import torch

from fira_testlib import load_batch_golden, load_raw_golden, reference_args


def _write_dataset(root, raw, n=None):
    os.makedirs(os.path.join(root, "DataSet"), exist_ok=True)
    for k, v in raw["raw"].items():
        json.dump(v[:n] if n else v, open(os.path.join(root, "DataSet", k + ".json"), "w"))
    json.dump(raw["word_vocab"], open(os.path.join(root, "DataSet", "word_vocab.json"), "w"))
    json.dump(raw["ast_change_vocab"], open(os.path.join(root, "DataSet", "ast_change_vocab.json"), "w"))
    json.dump(raw["VOCAB_UPPER_CASE"], open(os.path.join(root, "VOCAB_UPPER_CASE"), "w"))


def test_build_commit_equals_reference_process_data():
    from fira_icse_b200.data import build_commit
    raw = load_raw_golden()
    g = load_batch_golden()
    ptr = g["edge_ptr"]
    upper = set(raw["VOCAB_UPPER_CASE"])
    for i in range(128):
        c = build_commit(raw["raw"], i, raw["word_vocab"], raw["ast_change_vocab"], upper)
        for k in ("sou", "tar", "mark", "ast_change", "tar_label", "sub_token"):
            assert np.array_equal(np.array(c[k]), g[k][i]), (i, k)
        # adjacency: same set of entries, bit-identical float64 values
        ref = np.zeros((650, 650)); ref[g["edge_row"][ptr[i]:ptr[i + 1]], g["edge_col"][ptr[i]:ptr[i + 1]]] = \
            g["edge_val"][ptr[i]:ptr[i + 1]]
        mine = np.zeros((650, 650)); mine[np.repeat(np.arange(650), c["deg"]), c["col"]] = c["val"]
        assert np.array_equal(ref, mine), i
        attr = np.zeros((210, 25), np.int64)
        if c["attr_pos"]:
            attr[c["attr_pos"]] = c["attr_ids"]
        assert np.array_equal(attr, g["attr"][i]), (i, "attr")


def test_builders_equal_reference_on_extreme_and_truncated_commits():
    """tests/golden/make_golden_edge.py: the DataSet's extremes (longest diff, most AST/edit/sub-token nodes, commits
    without AST nodes / edit nodes / sub-tokens) and crafted commits that reach the truncation branches
    (diff > 208 tokens with edges into the cut-off part, message > 28 tokens) -- product builder (native adjacency)
    and oracle builder against the reference's own process_data output."""
    import graph_oracle as G
    from fira_testlib import load_edge_golden
    from fira_icse_b200.data import build_commit
    edge, g = load_edge_golden()
    base = load_raw_golden()
    vocab, ast_vocab, upper = base["word_vocab"], base["ast_change_vocab"], base["VOCAB_UPPER_CASE"]
    n = len(edge["notes"])
    assert n == len(g["sou"]) >= 14 and any("crafted" in s for s in edge["notes"])
    ptr = g["edge_ptr"]
    for i in range(n):
        c = build_commit(edge["raw"], i, vocab, ast_vocab, set(upper))
        o = G.build_commit(edge["raw"], i, vocab, ast_vocab, upper)
        ref = np.zeros((650, 650))
        ref[g["edge_row"][ptr[i]:ptr[i + 1]], g["edge_col"][ptr[i]:ptr[i + 1]]] = g["edge_val"][ptr[i]:ptr[i + 1]]
        mine = np.zeros((650, 650)); mine[np.repeat(np.arange(650), c["deg"]), c["col"]] = c["val"]
        orac = np.zeros((650, 650)); orac[o["row"], o["col"]] = o["val"]
        assert np.array_equal(ref, mine) and np.array_equal(ref, orac), edge["notes"][i]
        for k in ("sou", "tar", "mark", "ast_change", "tar_label", "sub_token"):
            assert np.array_equal(np.array(c[k]), g[k][i]), (edge["notes"][i], k)
            assert np.array_equal(np.array(o[k]), g[k][i]), (edge["notes"][i], k, "oracle")
        attr = np.zeros((210, 25), np.int64)
        if c["attr_pos"]:
            attr[c["attr_pos"]] = c["attr_ids"]
        assert np.array_equal(attr, g["attr"][i]), (edge["notes"][i], "attr")
    # the truncated diff really lost edges: no entry may touch a code node beyond the padded length
    j = next(k for k, s in enumerate(edge["notes"]) if "230 tokens" in s)
    assert len(edge["raw"]["difftoken"][j]) == 230 and (g["sou"][j] != 0).all()


def test_dataset_roundtrip_and_collate(tmp_path):
    from fira_icse_b200 import PackedEdges
    from fira_icse_b200.data import TransDataset, collate_packed
    raw = load_raw_golden()
    _write_dataset(str(tmp_path), raw, n=40)
    ds = TransDataset(reference_args(), "train", root=str(tmp_path))
    order = json.load(open(tmp_path / "all_index"))
    assert sorted(order["train"] + order["valid"] + order["test"]) == list(range(40))
    assert len(ds) == len(order["train"])
    g = load_batch_golden()
    i0 = order["train"][0]
    item = ds[0]
    assert np.array_equal(item[0], g["sou"][i0]) and np.array_equal(item[6], g["tar_label"][i0])
    assert np.array_equal(item[2], g["attr"][i0])
    batch = collate_packed([ds[i] for i in range(4)])
    assert batch[2] is None and batch[0].shape == (4, 210) and batch[0].dtype == torch.int64
    rowptr, col, val = batch[5]
    pe = PackedEdges(rowptr, col, val, 4, 650, True)
    dense = pe.to_dense(torch.float32)
    for b in range(4):
        i = order["train"][b]
        ptr = g["edge_ptr"]
        ref = torch.zeros(650, 650, dtype=torch.float64)
        ref[torch.from_numpy(g["edge_row"][ptr[i]:ptr[i + 1]].astype(np.int64)),
            torch.from_numpy(g["edge_col"][ptr[i]:ptr[i + 1]].astype(np.int64))] = torch.from_numpy(g["edge_val"][ptr[i]:ptr[i + 1]])
        assert torch.equal(dense[b], ref.float())
    dsd = TransDataset(reference_args(), "train", root=str(tmp_path), dense_edges=True)
    assert dsd[1][5].shape == (650, 650) and dsd[1][5].dtype == np.float64


def test_bleu_method2_known_values():
    from fira_icse_b200.bleu import sentence_bleu_method2 as bleu
    ref = "the cat sat on the mat".split()
    assert abs(bleu([ref], ref) - 1.0) < 1e-12
    assert bleu([ref], []) == 0.0
    assert bleu([ref], "dog barks".split()) == 0.0
    # hand computation (nltk floors an absent order's denominator at 1, then method2 adds 1/1):
    # hyp 'the cat sat' -> p1=3/3, p2=(2+1)/(2+1), p3=(1+1)/(1+1), p4=(0+1)/(1+1); BP=exp(1-6/3)
    import math
    assert abs(bleu([ref], "the cat sat".split()) - math.exp(-1.0) * 0.5 ** 0.25) < 1e-12
    # 'the the cat': p1 = 3/3 ('the' occurs twice in the reference), p2 = (1+1)/(2+1), p3 = p4 = (0+1)/(1+1)
    exp = math.exp(1 - 6 / 3) * math.exp(0.25 * (math.log(1.0) + math.log(2 / 3) + math.log(1 / 2) + math.log(1 / 2)))
    assert abs(bleu([ref], "the the cat".split()) - exp) < 1e-12


def test_native_adjacency_equals_oracle_builder_and_rejects_bad_input():
    """fira_host_build_adjacency (C++) against the pure-Python restatement of Dataset.py in oracle/graph_oracle.py
    (the comparison with the reference's own output is test_build_commit_equals_reference_process_data)."""
    import graph_oracle as G
    from fira_icse_b200._lib import FiraLibraryError
    from fira_icse_b200.data import build_adjacency, build_commit
    raw = load_raw_golden()
    upper = set(raw["VOCAB_UPPER_CASE"])
    for i in (0, 7, 63, 127):
        mine = build_commit(raw["raw"], i, raw["word_vocab"], raw["ast_change_vocab"], upper)
        ref = G.build_commit(raw["raw"], i, raw["word_vocab"], raw["ast_change_vocab"], raw["VOCAB_UPPER_CASE"])
        dense = np.zeros((650, 650)); dense[np.repeat(np.arange(650), mine["deg"]), mine["col"]] = mine["val"]
        want = np.zeros((650, 650)); want[ref["row"], ref["col"]] = ref["val"]
        assert np.array_equal(dense, want), i                   # float64 values bit-identical
        assert mine["deg"].sum() == len(mine["col"]) == len(mine["val"]) == len(ref["val"])
    # tiny hand-checked graph: code chain <start>-t1-<eos> (n_diff = 1) + one AST node tied to t1
    deg, col, val = build_adjacency([], [], [(0, 0)], [], [], n_diff=1, n_ast=1, diff_len=4, sub_len=2, ast_change_len=2)
    assert deg.tolist() == [2, 4, 2, 1, 1, 1, 2, 1]            # node 6 = first AST node, tied to code node 1
    assert col.tolist() == [0, 1, 0, 1, 2, 6, 1, 2, 3, 4, 5, 1, 6, 7]
    import math
    assert val[0] == 1 / math.sqrt(2) / math.sqrt(2) and val[1] == 1 / math.sqrt(2) / math.sqrt(4) and val[-1] == 1.0
    with pytest.raises(FiraLibraryError):                       # AST id outside the padded graph
        build_adjacency([], [], [], [(0, 5)], [], n_diff=1, n_ast=1, diff_len=4, sub_len=2, ast_change_len=2)
    with pytest.raises(FiraLibraryError):                       # self edge
        build_adjacency([], [], [], [(1, 1)], [], n_diff=1, n_ast=2, diff_len=4, sub_len=2, ast_change_len=2)
    # edges to code tokens beyond the padded diff are dropped, not an error (Dataset.py:228,243)
    deg2, _, _ = build_adjacency([], [], [(0, 3)], [], [], n_diff=1, n_ast=1, diff_len=4, sub_len=2, ast_change_len=2)
    assert deg2.tolist() == [2, 3, 2, 1, 1, 1, 1, 1]


def test_native_loader_equals_python_collate_and_trim(tmp_path):
    """PackedBatchLoader (fira_host_batch_dims + fira_host_gather_batch) == collate_packed + trim_batch_host."""
    from fira_icse_b200.data import PackedBatchLoader, TransDataset, collate_packed, trim_batch_host
    raw = load_raw_golden()
    _write_dataset(str(tmp_path), raw, n=60)
    ds = TransDataset(reference_args(), "train", root=str(tmp_path))
    V = len(raw["word_vocab"])
    n = len(ds)
    ld = PackedBatchLoader(ds, 8, V, shuffle=False, multiples=(8, 8, 8), pin=False)
    assert len(ld) == -(-n // 8)
    seen = 0
    for k, got in enumerate(ld):
        idx = list(range(k * 8, min(n, k * 8 + 8)))
        want = trim_batch_host(collate_packed([ds[i] for i in idx]), V)
        for j in (0, 1, 3, 4, 6, 7):
            assert got[j].dtype == torch.int64 and torch.equal(got[j], want[j]), (k, j)
        assert got[2] is None
        for a, b in zip(got[5], want[5]):
            assert a.dtype == b.dtype and torch.equal(a, b), k
        seen += len(idx)
    assert seen == n
    # untrimmed mode reproduces collate_packed itself
    full = PackedBatchLoader(ds, 5, V, multiples=None, pin=False).gather(np.arange(5))
    want = collate_packed([ds[i] for i in range(5)])
    assert all(torch.equal(full[j], want[j]) for j in (0, 1, 3, 4, 6, 7))
    assert all(torch.equal(a, b) for a, b in zip(full[5], want[5]))
    # shape budget: with one allowed shape every later batch is padded up to a shape that holds it
    capped = PackedBatchLoader(ds, 8, V, multiples=(8, 8, 8), max_shapes=1, pin=False)
    shapes = [(b[0].shape[1], b[7].shape[1], b[4].shape[1]) for b in capped]
    assert len(set(shapes)) <= 2 and all(s == shapes[0] or s == (210, 160, 280) for s in shapes)
    # shuffled epochs are permutations driven by torch's generator
    torch.manual_seed(3)
    sh = PackedBatchLoader(ds, 8, V, shuffle=True, pin=False)
    first = torch.cat([b[1].clone() for b in sh])          # batches are views of recycled staging slots
    torch.manual_seed(3)
    again = torch.cat([b[1].clone() for b in sh])
    assert torch.equal(first, again) and first.shape[0] == n
    assert not torch.equal(first, torch.cat([b[1].clone() for b in ld]))


def test_bucketed_batching_visits_every_commit_once_and_trims_more(tmp_path):
    """opt-in size-bucketed batches (PackedBatchLoader(bucket=K)): a permutation of the epoch, far fewer padded rows"""
    from fira_icse_b200.data import PackedBatchLoader, TransDataset
    raw = load_raw_golden()
    _write_dataset(str(tmp_path), raw)
    ds = TransDataset(reference_args(), "train", root=str(tmp_path))
    V, n = len(raw["word_vocab"]), len(ds)

    def epoch(loader):
        seen, rows = [], 0
        for b in loader:
            seen.append(b[1].clone())                             # views of recycled staging slots
            rows += b[0].shape[0] * (b[0].shape[1] + b[7].shape[1] + b[4].shape[1])
        return torch.cat(seen), rows
    torch.manual_seed(11)
    plain_tar, plain_rows = epoch(PackedBatchLoader(ds, 8, V, shuffle=True, pin=False))
    torch.manual_seed(11)
    ld = PackedBatchLoader(ds, 8, V, shuffle=True, pin=False, bucket=4)
    buck_tar, buck_rows = epoch(ld)
    assert buck_tar.shape[0] == plain_tar.shape[0] == n
    key = lambda t: sorted(map(tuple, t.tolist()))
    assert key(buck_tar) == key(plain_tar)                       # same multiset of commits, each exactly once
    assert buck_rows < 0.85 * plain_rows, (buck_rows, plain_rows)
    sizes = [len(c) for c in ld.epoch_batches()]
    assert sum(sizes) == n and sizes.count(8) == n // 8 and all(s == 8 for s in sizes[:n // 8])
    assert all(len(c) == 8 for c in PackedBatchLoader(ds, 8, V, shuffle=True, pin=False, bucket=4,
                                                      drop_last=True).epoch_batches())
    torch.manual_seed(11)
    again, _ = epoch(ld)
    assert torch.equal(again, buck_tar)                           # reproducible under torch.manual_seed


def test_trim_batch_host_keeps_real_rows_and_renumbers():
    """padding trimming: shapes shrink to the batch maximum, adjacency of real nodes and labels stay equivalent"""
    from fira_icse_b200 import PackedEdges
    from fira_icse_b200.data import collate_packed, trim_batch_host, build_commit
    raw = load_raw_golden()
    upper = set(raw["VOCAB_UPPER_CASE"])
    V = len(raw["word_vocab"])
    items = []
    for i in range(6):
        c = build_commit(raw["raw"], i, raw["word_vocab"], raw["ast_change_vocab"], upper)
        items.append([np.array(c["sou"]), np.array(c["tar"]), None, np.array(c["mark"]), np.array(c["ast_change"]),
                      (c["deg"], c["col"], c["val"]), np.array(c["tar_label"]), np.array(c["sub_token"])])
    full = collate_packed(items)
    trim = trim_batch_host(full, V)
    c0, c1, c2 = trim[0].shape[1], trim[7].shape[1], trim[4].shape[1]
    assert c0 < 210 and c1 < 160 and c2 < 280 and c0 % 8 == 0
    assert torch.equal(trim[0], full[0][:, :c0]) and (full[0][:, c0:] == 0).all()
    assert (full[7][:, c1:] == 0).all() and (full[4][:, c2:] == 0).all()
    dense_full = PackedEdges(*full[5], 6, 650, True).to_dense()
    dense_trim = PackedEdges(*trim[5], 6, c0 + c1 + c2, True).to_dense()
    keep = np.r_[0:c0, 210:210 + c1, 370:370 + c2]
    assert torch.equal(dense_trim, dense_full[:, keep][:, :, keep])
    # removed rows were isolated self loops
    gone = np.setdiff1d(np.arange(650), keep)
    assert torch.equal(dense_full[:, gone][:, :, gone], torch.eye(len(gone), dtype=torch.float64).expand(6, -1, -1))
    # labels: vocabulary and code-copy labels unchanged, sub-token copy labels shifted by the trimmed code padding
    lf, lt = full[6], trim[6]
    sub = lf >= V + 210
    assert torch.equal(lt[~sub], lf[~sub]) and torch.equal(lt[sub], lf[sub] - (210 - c0))


def test_synth_dataset_through_native_loader_equals_bench_host_batches():
    """bench.py's `e2e_loader` arm: SynthDataset served by PackedBatchLoader == the batches bench.py builds in Python"""
    import bench
    from fira_icse_b200.data import PackedBatchLoader
    from fira_icse_b200.synth import SynthDataset
    ds = SynthDataset(100, 12, bench.VOCAB, bench.AST_VOCAB)
    ld = PackedBatchLoader(ds, 6, bench.VOCAB, shuffle=False, multiples=(8, 8, 8), pin=False)
    n = 0
    for k, got in enumerate(ld):
        t, csr, _ = bench.host_batch(100 + 6 * k, 6, pin=False, trim=True)
        for j, key in ((0, "sou"), (1, "tar"), (3, "mark"), (4, "ast_change"), (6, "tar_label"), (7, "sub_token")):
            assert torch.equal(got[j], t[key]), (k, key)
        for a, b in zip(got[5], csr):
            assert a.dtype == b.dtype and torch.equal(a, b), k
        n += 1
    assert n == 2
