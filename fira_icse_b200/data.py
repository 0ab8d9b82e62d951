"""`Dataset.py` surface of the reference (TransDataset) with a packed on-disk/in-memory format.

Same inputs (CWD-relative DataSet/*.json, VOCAB_UPPER_CASE, all_index), same id conversion,
padding, dual-copy labels and adjacency definition as the reference (Dataset.py:96-294,346-357);
what changes is the representation handed to the model:

  reference: per item a dense float64 650x650 `toarray()` (3.38 MB/commit through collate + PCIe)
  here     : per item the CSR pieces (row degrees uint8[650], col int16[nnz], val fp32[nnz], ~6 KB);
             `collate_packed` concatenates them into one batch CSR (pinned host tensors) that
             PackedEdges.from_host ships with three async H2D copies.

`TransDataset[i]` still returns the reference's 8-element list (dense adjacency on request) so the
reference's own DataLoader/collate keeps working.
"""
import json
import math
import os
import random

import numpy as np
import torch
from torch.utils.data import Dataset

num_train, num_valid, num_test = 75000, 8000, 7661          # Dataset.py:10-12
lemmatization = {"added": "add", "fixed": "fix", "removed": "remove",
                 "adding": "add", "fixing": "fix", "removing": "remove"}
RAW_FILES = ["difftoken", "diffatt", "diffmark", "msg", "variable", "change", "ast",
             "edge_change_code", "edge_change_ast", "edge_ast_code", "edge_ast"]


def _lower(tok, upper):
    return tok if tok in upper else tok.lower()


def _to_ids(tokens, vocab, upper):
    out = []
    for t in tokens:
        t = _lower(t, upper)
        out.append(vocab[t] if t in vocab else vocab["<unkm>"])     # lazy like Dataset.py:75-78
    return out


def _fit(seq, n):
    return (list(seq) + [0] * n)[:n]


def build_commit(raw, i, vocab, ast_vocab, upper, diff_len=210, msg_len=30, att_len=25, ast_change_len=280,
                 sub_len=160):
    """One commit -> padded id arrays + CSR pieces of its normalised adjacency.

    Node ids: code token j -> j+1 (0 = <start>), sub-token k -> 210+k, AST node a -> 370+a,
    edit node c -> 370+len(ast)+c.  Edges: edit-code, edit-AST, AST-code, AST-AST, code-sub-token,
    sequential code chain; undirected, de-duplicated; self loop on all 650 nodes;
    value 1/sqrt(deg_row)/sqrt(deg_col) in float64."""
    var_map = raw["variable"][i]
    diff = [_lower(var_map.get(t, t), upper) for t in raw["difftoken"][i]]
    msg = [lemmatization.get(w, w) for w in (_lower(var_map.get(t, t), upper) for t in raw["msg"][i])]
    atts = raw["diffatt"][i]
    V = len(vocab)
    n_ast = len(raw["ast"][i])
    n_nodes = diff_len + sub_len + ast_change_len

    sou = _fit([vocab["<start>"]] + _to_ids(diff, vocab, upper) + [vocab["<eos>"]], diff_len)
    msg_ids = _to_ids(msg, vocab, upper)
    tar = _fit([vocab["<start>"]] + msg_ids + [vocab["<eos>"]], msg_len)
    mark = _fit([2] + list(raw["diffmark"][i]) + [2], diff_len)
    ast_change = _fit(_to_ids(list(raw["ast"][i]) + list(raw["change"][i]), ast_vocab, upper), ast_change_len)

    # sub-token nodes are shared by repeated identifiers (first occurrence defines them)
    sub_tokens, owner, code_sub = [], {}, []
    for j, att in enumerate(atts):
        if att:
            if diff[j] not in owner:
                owner[diff[j]] = range(len(sub_tokens), len(sub_tokens) + len(att))
                sub_tokens.extend(att)
            code_sub.extend((j, k) for k in owner[diff[j]])
    sub_token = _fit(_to_ids(sub_tokens, vocab, upper), sub_len)

    # dual-copy labels: position in the diff wins over position among the sub-tokens
    first_in_diff, first_in_sub = {}, {}
    for j, t in enumerate(diff):
        first_in_diff.setdefault(t, j)
    for k, t in enumerate(sub_tokens):
        first_in_sub.setdefault(t, k)
    label = []
    for w, wid in zip(msg, msg_ids):
        if w in first_in_diff:
            label.append(first_in_diff[w] + V + 1)
        elif w in first_in_sub:
            label.append(first_in_sub[w] + V + diff_len)
        else:
            label.append(wid)
    tar_label = _fit([vocab["<start>"]] + label + [vocab["<eos>"]], msg_len)

    # adjacency as a set of ordered pairs, keyed r*n + c
    a0 = diff_len + sub_len
    und = []
    und += [(c + a0 + n_ast, j + 1) for c, j in raw["edge_change_code"][i] if j + 1 < diff_len]
    und += [(c + a0 + n_ast, a + a0) for c, a in raw["edge_change_ast"][i]]
    und += [(a + a0, j + 1) for a, j in raw["edge_ast_code"][i] if j + 1 < diff_len]
    und += [(a + a0, b + a0) for a, b in raw["edge_ast"][i]]
    und += [(j + 1, k + diff_len) for j, k in code_sub]
    und += [(j, j + 1) for j in range(len(diff) + 1)]
    e = np.array(und, np.int64).reshape(-1, 2)
    assert (e[:, 0] != e[:, 1]).all(), "the DataSet has no self edges (Dataset.py:275)"
    keys = np.unique(np.concatenate((e[:, 0] * n_nodes + e[:, 1], e[:, 1] * n_nodes + e[:, 0],
                                     np.arange(n_nodes) * (n_nodes + 1))))
    row, col = keys // n_nodes, keys % n_nodes
    deg_r = np.bincount(row, minlength=n_nodes)
    deg_c = np.bincount(col, minlength=n_nodes)
    val = np.array([1 / math.sqrt(deg_r[r]) / math.sqrt(deg_c[c]) for r, c in zip(row, col)], np.float64)
    attr_pos = [j + 1 for j, att in enumerate(atts) if att]           # row of the padded [210,25] attr matrix
    attr_ids = [_fit(_to_ids(atts[j - 1], vocab, upper), att_len) for j in attr_pos]
    return dict(sou=sou, tar=tar, mark=mark, ast_change=ast_change, tar_label=tar_label, sub_token=sub_token,
                deg=deg_r.astype(np.uint8), col=col.astype(np.int16), val=val,
                attr_pos=[p for p in attr_pos if p < diff_len], attr_ids=attr_ids[:sum(p < diff_len for p in attr_pos)])


class TransDataset(Dataset):
    """TransDataset(args, 'train'|'valid'|'test'[, root='.'])  (Dataset.py:17-68)."""

    ID_KEYS = ("sou", "tar", "mark", "ast_change", "tar_label", "sub_token")

    def __init__(self, args, data_name, root=".", dense_edges=False, limit=None):
        super().__init__()
        self.data_name = data_name
        self.diff_len, self.msg_len, self.att_len = args.sou_len, args.tar_len, args.att_len
        self.ast_change_len, self.sub_token_len = args.ast_change_len, args.sub_token_len
        self.graph_len = self.diff_len + self.sub_token_len + self.ast_change_len
        self.dense_edges = dense_edges
        self.root = root
        cache = os.path.join(root, f"processed_b200_{data_name}.npz")
        if not os.path.exists(cache):
            self._process_all(limit)
        z = np.load(cache)
        self.d = {k: z[k] for k in z.files}
        print("Loaded data!")

    # ------------------------------------------------------------------ one-off preprocessing
    def _process_all(self, limit=None):
        j = lambda n: json.load(open(os.path.join(self.root, "DataSet", n + ".json")))
        raw = {n: j(n) for n in RAW_FILES}
        n_all = len(raw["difftoken"])
        assert all(len(raw[n]) == n_all for n in RAW_FILES)
        vocab, ast_vocab = j("word_vocab"), j("ast_change_vocab")
        upper = set(json.load(open(os.path.join(self.root, "VOCAB_UPPER_CASE"))))
        idx_path = os.path.join(self.root, "all_index")
        if os.path.exists(idx_path):
            all_index = json.load(open(idx_path))
        else:                                      # Dataset.py:306-313 under seed_everything(0)
            index = list(range(n_all))
            random.Random(0).shuffle(index)
            nt, nv = min(num_train, int(n_all * 0.83)), min(num_valid, int(n_all * 0.09))
            all_index = {"train": index[:nt], "valid": index[nt:nt + nv], "test": index[nt + nv:]}
            json.dump(all_index, open(idx_path, "w"))
        for split, order in all_index.items():
            if limit:
                order = order[:limit]
            cols = {k: [] for k in self.ID_KEYS}
            deg, col, val, eptr = [], [], [], [0]
            apos, aids, aptr = [], [], [0]
            for i in order:
                c = build_commit(raw, i, vocab, ast_vocab, upper, self.diff_len, self.msg_len, self.att_len,
                                 self.ast_change_len, self.sub_token_len)
                for k in self.ID_KEYS:
                    cols[k].append(c[k])
                deg.append(c["deg"]); col.append(c["col"]); val.append(c["val"]); eptr.append(eptr[-1] + len(c["col"]))
                apos += c["attr_pos"]; aids += c["attr_ids"]; aptr.append(aptr[-1] + len(c["attr_pos"]))
            np.savez(os.path.join(self.root, f"processed_b200_{split}.npz"),
                     **{k: np.array(v, np.int32).reshape(len(order), -1) for k, v in cols.items()},
                     deg=np.array(deg, np.uint8).reshape(len(order), self.graph_len),
                     col=np.concatenate(col) if col else np.zeros(0, np.int16),
                     val=np.concatenate(val) if val else np.zeros(0, np.float64), edge_ptr=np.array(eptr, np.int64),
                     attr_pos=np.array(apos, np.int16), attr_ids=np.array(aids, np.int32).reshape(-1, self.att_len),
                     attr_ptr=np.array(aptr, np.int64), index=np.array(order, np.int64))

    # ------------------------------------------------------------------ access
    def __len__(self):
        return len(self.d["sou"])

    def csr_pieces(self, i):
        lo, hi = self.d["edge_ptr"][i], self.d["edge_ptr"][i + 1]
        return self.d["deg"][i], self.d["col"][lo:hi], self.d["val"][lo:hi]

    def attr(self, i):
        out = np.zeros((self.diff_len, self.att_len), np.int64)
        lo, hi = self.d["attr_ptr"][i], self.d["attr_ptr"][i + 1]
        out[self.d["attr_pos"][lo:hi]] = self.d["attr_ids"][lo:hi]
        return out

    def dense_edge(self, i):
        deg, col, val = self.csr_pieces(i)
        a = np.zeros((self.graph_len, self.graph_len), np.float64)
        a[np.repeat(np.arange(self.graph_len), deg), col] = val
        return a

    def __getitem__(self, i):
        d = self.d
        edge = self.dense_edge(i) if self.dense_edges else self.csr_pieces(i)
        return [d["sou"][i].astype(np.int64), d["tar"][i].astype(np.int64), self.attr(i), d["mark"][i].astype(np.int64),
                d["ast_change"][i].astype(np.int64), edge, d["tar_label"][i].astype(np.int64),
                d["sub_token"][i].astype(np.int64)]


def collate_packed(items, pin=False):
    """list of TransDataset items (csr pieces) -> the 8-element batch with item 5 = host CSR triple
    (rowptr int32 [B*650+1], col int32, val fp32).  `attr` is dropped (None): the model ignores it."""
    def stack(k):
        return torch.from_numpy(np.stack([it[k] for it in items]))
    deg = np.concatenate([it[5][0] for it in items]).astype(np.int64)
    rowptr = torch.from_numpy(np.concatenate((np.zeros(1, np.int64), np.cumsum(deg))).astype(np.int32))
    col = torch.from_numpy(np.concatenate([it[5][1] for it in items]).astype(np.int32))
    val = torch.from_numpy(np.concatenate([it[5][2] for it in items]).astype(np.float32))
    out = [stack(0), stack(1), None, stack(3), stack(4), (rowptr, col, val), stack(6), stack(7)]
    if pin:
        out = [tuple(x.pin_memory() for x in o) if isinstance(o, tuple) else (o.pin_memory() if o is not None else None)
               for o in out]
    return out


def batch_to_device(batch, device, n_nodes=650):
    """collate_packed output -> model inputs on `device` (async copies from pinned memory)."""
    from .graph import PackedEdges
    out = []
    for j, o in enumerate(batch):
        if j == 5:
            B = batch[0].shape[0]
            out.append(PackedEdges.from_host(*o, B, n_nodes, device))
        else:
            out.append(o.to(device, non_blocking=True) if o is not None else None)
    return out


def trim_batch_host(batch, vocab_size, multiple=8, full=(210, 160, 280)):
    """Drop the padding the whole batch shares (var-len packing, SURVEY.md section 8f rank 4, loader side).

    batch: collate_packed-style list [sou, tar, attr, mark, ast_change, (rowptr, col, val), tar_label, sub_token]
    (host tensors, node lengths `full`).  Code / sub-token / AST+edit segments are cut to the longest
    commit of the batch (rounded up to `multiple`); the adjacency rows of the removed nodes -- isolated
    self-loop rows by construction (Dataset.py:271-275) -- are dropped, column ids and sub-token copy
    labels are renumbered.  Real rows, loss and gradients are unchanged; the model takes the shorter
    tensors as they are (all shapes are read from the inputs)."""
    sou, tar, attr, mark, ast_change, (rowptr, col, val), tar_label, sub_token = batch
    n0, n1, n2 = full
    B = sou.shape[0]

    def cap(t, n):
        used = int((t != 0).sum(1).max()) if t.numel() else 0
        return min(n, max(multiple, -(-used // multiple) * multiple))
    c0, c1, c2 = cap(sou, n0), cap(sub_token, n1), cap(ast_change, n2)
    N, Nt = n0 + n1 + n2, c0 + c1 + c2
    keep = np.zeros(N, bool)
    keep[:c0] = True; keep[n0:n0 + c1] = True; keep[n0 + n1:n0 + n1 + c2] = True
    remap = np.full(N, -1, np.int64)
    remap[keep] = np.arange(Nt)
    rp = rowptr.numpy().astype(np.int64)
    deg = np.diff(rp).reshape(B, N)
    row_keep = np.broadcast_to(keep, (B, N)).reshape(-1)
    entry_keep = np.repeat(row_keep, deg.reshape(-1))
    new_col = remap[col.numpy()[entry_keep]]
    assert (new_col >= 0).all(), "a kept node has a neighbour inside the trimmed padding: adjacency is not padding-isolated"
    new_rowptr = np.concatenate((np.zeros(1, np.int64), np.cumsum(deg[:, keep].reshape(-1))))
    label = tar_label.clone()
    sub_copy = label >= vocab_size + n0
    label[sub_copy] -= (n0 - c0)
    out = [sou[:, :c0].contiguous(), tar, attr, mark[:, :c0].contiguous(), ast_change[:, :c2].contiguous(),
           (torch.from_numpy(new_rowptr.astype(np.int32)), torch.from_numpy(new_col.astype(np.int32)),
            torch.from_numpy(val.numpy()[entry_keep])), label, sub_token[:, :c1].contiguous()]
    return out
