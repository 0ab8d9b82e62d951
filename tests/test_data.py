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
