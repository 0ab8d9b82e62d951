#!/usr/bin/env python
"""Edge-case fixture for the graph/batch builder, produced by the UNMODIFIED reference Dataset.process_data.

The first-128 fixture (make_golden.py) holds ordinary commits.  This one holds
  * the extremes of the full 90,661-commit DataSet: longest diff (198 tokens), most AST nodes (90), most edit
    nodes (99), most AST+edit nodes (157), most sub-tokens (256 raw), shortest diff (9), commits without AST
    nodes / edit nodes / sub-tokens, the highest code index referenced by an edge (194);
  * crafted commits that exercise the truncation branches the shipped data never reaches (Dataset.py:141-171 cut
    sequences to 210 / 30 / 280 / 160, Dataset.py:228,243 drop edges to code tokens beyond the padded diff): a real
    commit whose diff is repeated until it is longer than 208 tokens, with AST/edit edges pointing into the cut-off
    part, and one whose message is longer than 28 tokens.  Crafted commits the reference itself cannot process
    are reported and left out.

Runs only in the build container (needs /root/reference).  Writes tests/golden/raw_edge.json.gz and
tests/golden/batch_edge.npz (same layout as batch_first128.npz).

Usage:  python tests/golden/make_golden_edge.py
"""
import copy
import gzip
import json
import os
import pickle
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
RAW_NAMES = ["difftoken", "diffatt", "diffmark", "msg", "variable", "change", "ast",
             "edge_change_code", "edge_change_ast", "edge_ast_code", "edge_ast"]


def pick_extremes(full):
    n = len(full["difftoken"])
    nd = np.array([len(x) for x in full["difftoken"]])
    nm = np.array([len(x) for x in full["msg"]])
    na = np.array([len(x) for x in full["ast"]])
    nc = np.array([len(x) for x in full["change"]])
    ns = np.array([sum(len(a) for a in x) for x in full["diffatt"]])
    hi = np.array([max([e[1] for e in a] + [e[1] for e in b] + [-1])
                   for a, b in zip(full["edge_change_code"], full["edge_ast_code"])])
    ne = np.array([sum(len(full[k][i]) for k in RAW_NAMES[7:]) for i in range(n)])
    sel = [int(nd.argmax()), int(nd.argmin()), int(nm.argmax()), int(nm.argmin()), int(na.argmax()), int(nc.argmax()),
           int((na + nc).argmax()), int(ns.argmax()), int(hi.argmax()), int(ne.argmax()), int(ne.argmin())]
    sel += [int(i) for i in np.flatnonzero(na == 0)[:2]]
    sel += [int(i) for i in np.flatnonzero((na == 0) & (nc == 0))[:2]]
    sel += [int(i) for i in np.flatnonzero(ns == 0)[:2]]
    sel += [int(i) for i in np.flatnonzero((ns == 0) & (na + nc == 0))[:1]]
    out = []
    for i in sel:
        if i not in out:
            out.append(i)
    return out


def crafted(full, base):
    """Variations of commit `base` that reach the truncation branches."""
    c = {k: copy.deepcopy(full[k][base]) for k in RAW_NAMES}
    n0 = len(c["difftoken"])
    reps = 215 // n0 + 1
    long_diff = copy.deepcopy(c)
    for k in ("difftoken", "diffatt", "diffmark"):
        long_diff[k] = (c[k] * reps)[:230]
    # edges into the part that survives, the last kept position and the cut-off tail
    long_diff["edge_ast_code"] = c["edge_ast_code"] + [[0, 207], [0, 208], [0, 209], [0, 229]]
    long_diff["edge_change_code"] = c["edge_change_code"] + ([[0, 208], [0, 215]] if c["change"] else [])
    long_msg = copy.deepcopy(c)
    long_msg["msg"] = (c["msg"] * 12)[:35]
    return [("diff of 230 tokens with edges beyond position 208", long_diff),
            ("message of 35 tokens", long_msg)]


def run_reference(raw, word_vocab, ast_vocab):
    import Dataset as RefDataset
    n = len(raw["difftoken"])
    RefDataset.num_train, RefDataset.num_valid, RefDataset.num_test = n, 0, 0
    ds = object.__new__(RefDataset.TransDataset)
    ds.data_name = "train"
    ds.diff_len, ds.msg_len, ds.att_len = 210, 30, 25
    ds.ast_change_len, ds.sub_token_len = 280, 160
    ds.graph_len = 650
    ds.vocab, ds.ast_change_vocab = word_vocab, ast_vocab
    for f in ("processed_train.pkl", "all_index"):
        if os.path.exists(f):
            os.remove(f)
    r = json.loads(json.dumps(raw))          # process_data mutates its inputs
    ds.process_data(*[r[k] for k in RAW_NAMES])
    data = pickle.load(open("processed_train.pkl", "rb"))
    order = json.load(open("all_index"))["train"]
    inv = np.argsort(np.array(order))
    return [np.asarray(data[i])[inv] if i != 5 else [data[5][j] for j in inv] for i in range(8)]


def main():
    scratch = tempfile.mkdtemp(prefix="fira_golden_edge_")
    os.symlink(os.path.join(REF, "DataSet"), os.path.join(scratch, "DataSet"))
    os.symlink(os.path.join(REF, "VOCAB_UPPER_CASE"), os.path.join(scratch, "VOCAB_UPPER_CASE"))
    os.chdir(scratch)
    sys.path.insert(0, REF)
    full = {k: json.load(open(os.path.join(REF, "DataSet", k + ".json"))) for k in RAW_NAMES}
    word_vocab = json.load(open(os.path.join(REF, "DataSet", "word_vocab.json")))
    ast_vocab = json.load(open(os.path.join(REF, "DataSet", "ast_change_vocab.json")))
    upper = json.load(open(os.path.join(REF, "VOCAB_UPPER_CASE")))
    picked = pick_extremes(full)
    raw = {k: [copy.deepcopy(full[k][i]) for i in picked] for k in RAW_NAMES}
    notes = [f"DataSet commit {i}" for i in picked]
    base = int(np.argmax([len(full["edge_ast_code"][i]) > 4 and len(full["change"][i]) > 0 and
                          40 <= len(full["difftoken"][i]) <= 80 for i in range(2000)]))
    for what, c in crafted(full, base):
        trial = {k: raw[k] + [c[k]] for k in RAW_NAMES}
        try:
            run_reference(trial, word_vocab, ast_vocab)
        except Exception as exc:               # the reference cannot process it: not a parity case
            print(f"[edge] crafted commit '{what}' rejected by the reference: {type(exc).__name__}: {exc}")
            continue
        raw = trial
        notes.append(f"crafted from DataSet commit {base}: {what}")
    sou, tar, attr, mark, ast_change, edges, tar_label, sub_token = run_reference(raw, word_vocab, ast_vocab)
    ptr, rows, cols, vals = [0], [], [], []
    for e in edges:
        e = e.tocoo()
        rows.append(e.row.astype(np.int16)); cols.append(e.col.astype(np.int16))
        vals.append(e.data.astype(np.float64)); ptr.append(ptr[-1] + e.nnz)
    np.savez_compressed(os.path.join(HERE, "batch_edge.npz"),
                        sou=sou.astype(np.int16), tar=tar.astype(np.int16), attr=attr.astype(np.int16),
                        mark=mark.astype(np.int8), ast_change=ast_change.astype(np.int16),
                        tar_label=tar_label.astype(np.int16), sub_token=sub_token.astype(np.int16),
                        edge_ptr=np.array(ptr, np.int32), edge_row=np.concatenate(rows),
                        edge_col=np.concatenate(cols), edge_val=np.concatenate(vals))
    # ---- the reference model (torch.manual_seed(0) initialisation, eval mode) on these commits
    import torch
    import Model as RefModel
    torch.set_num_threads(8)

    class DotDict(dict):
        def __getattr__(self, k):
            return self[k]
    args = DotDict(sou_len=210, tar_len=30, att_len=25, ast_change_len=280, sub_token_len=160, lr=1e-4,
                   dropout_rate=0.1, num_head=8, embedding_dim=256, vocab_size=len(word_vocab),
                   ast_change_vocab_size=len(ast_vocab))
    torch.manual_seed(0)
    model = RefModel.TransModel(args)
    model.eval()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long()
    dense = torch.from_numpy(np.stack([e.toarray() for e in edges]))
    batch = [t(sou), t(tar), t(attr), t(mark), t(ast_change), dense, t(tar_label), t(sub_token)]
    with torch.no_grad():
        loss_sum, mask_sum = model(*batch, "train")
        ids = model(*batch, "dev")
        per_commit = [model(*[b[i:i + 1] for b in batch], "train")[0].item() for i in range(len(notes))]
    np.savez_compressed(os.path.join(HERE, "model_edge.npz"), loss_sum=loss_sum.item(), mask_sum=mask_sum.item(),
                        argmax_ids=ids.numpy().astype(np.int32), loss_per_commit=np.array(per_commit))
    with gzip.open(os.path.join(HERE, "raw_edge.json.gz"), "wt") as f:
        json.dump({"raw": raw, "notes": notes, "word_vocab_file": "raw_first128.json.gz (same vocabularies)"}, f)
    print(f"[edge] {len(notes)} commits written")
    for s in notes:
        print("   ", s)


if __name__ == "__main__":
    main()
