"""Shared helpers for the test-suite (args, golden loading, oracle <-> product glue)."""
import gzip
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


class DotDict(dict):
    def __getattr__(self, k):
        return self[k]


def reference_args(vocab_size=24650, ast_change_vocab_size=71):
    """run_model.py:27-56 hyper-parameters."""
    return DotDict(sou_len=210, tar_len=30, att_len=25, ast_change_len=280, sub_token_len=160, lr=1e-4,
                   dropout_rate=0.1, num_head=8, embedding_dim=256, vocab_size=vocab_size,
                   ast_change_vocab_size=ast_change_vocab_size)


def load_batch_golden():
    return np.load(os.path.join(GOLDEN, "batch_first128.npz"))


def load_model_golden():
    return np.load(os.path.join(GOLDEN, "model_first128.npz"), allow_pickle=False)


def load_raw_golden():
    with gzip.open(os.path.join(GOLDEN, "raw_first128.json.gz"), "rt") as f:
        return json.load(f)


def load_edge_golden():
    """(raw commits + notes, reference process_data output) of tests/golden/make_golden_edge.py: DataSet extremes and
    crafted commits that reach the truncation branches."""
    with gzip.open(os.path.join(GOLDEN, "raw_edge.json.gz"), "rt") as f:
        raw = json.load(f)
    return raw, np.load(os.path.join(GOLDEN, "batch_edge.npz"))


def golden_batch(lo, hi, dense_edge=True, edge_dtype=torch.float64):
    """The 8-tuple Dataset.__getitem__/collate would hand the model for commits [lo, hi)."""
    g = load_batch_golden()
    t = lambda k: torch.from_numpy(g[k][lo:hi].astype(np.int64))
    ptr = g["edge_ptr"]
    coo = [(g["edge_row"][ptr[i]:ptr[i + 1]].astype(np.int64), g["edge_col"][ptr[i]:ptr[i + 1]].astype(np.int64),
            g["edge_val"][ptr[i]:ptr[i + 1]]) for i in range(lo, hi)]
    if dense_edge:
        edge = torch.zeros(hi - lo, 650, 650, dtype=torch.float64)
        for b, (r, c, v) in enumerate(coo):
            edge[b].index_put_((torch.from_numpy(r), torch.from_numpy(c)), torch.from_numpy(v), accumulate=True)
        edge = edge.to(edge_dtype)
    else:
        edge = coo
    return [t("sou"), t("tar"), t("attr"), t("mark"), t("ast_change"), edge, t("tar_label"), t("sub_token")]


_MODEL_CACHE = {}


def seeded_model():
    """Product TransModel initialised exactly like the reference under torch.manual_seed(0) (CPU)."""
    if "m" not in _MODEL_CACHE:
        import fira_icse_b200 as F
        torch.manual_seed(0)
        _MODEL_CACHE["m"] = F.TransModel(reference_args())
    return _MODEL_CACHE["m"]
