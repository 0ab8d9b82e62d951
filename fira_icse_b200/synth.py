"""Synthetic commits with the DataSet's node/edge distribution (BASELINE.json:
"throughput on synthetic graphs matching the DataSet's node/edge distribution").

Per commit (all seeded by the commit index, so every rank/arm regenerates the same data):
  code tokens incl. <start>/<eos>   mean ~100, p90 ~165, max 200      (2 + Gamma(4, 24.6))
  sub-token nodes                   mean ~27,  p90 ~47,  max 102      (Gamma(3, 9))
  AST + edit nodes                  mean ~32,  p90 ~64,  max 157      (Gamma(2.2, 14.4), 80 % AST)
  message tokens                    mean ~6.8, p90 ~11,  max 20       (1 + Gamma(3.5, 1.66))
  relations (undirected, de-duplicated, both directions stored, Dataset.py:346-357):
     sequential code chain, code<->sub-token (~1.7 per sub-token), AST<->code (~1 per AST node),
     AST<->AST forest, edit<->code, edit<->AST (1 each)   -> ~400 directed off-diagonal entries
  + a self loop on all 650 nodes, value 1/sqrt(deg_r)/sqrt(deg_c) (Dataset.py:271-291)
  labels: 66 % vocabulary, 10 % copy-from-code, 12 % copy-from-sub-token, then <eos>
"""
import numpy as np

N_CODE, N_SUB, N_AST, T_LEN = 210, 160, 280, 30
N_NODES = N_CODE + N_SUB + N_AST
START, EOS = 1, 2          # vocabulary ids of <start>/<eos> in DataSet/word_vocab.json order are irrelevant here


def _clip_gamma(rng, shape, scale, lo, hi, offset=0):
    return int(np.clip(round(offset + rng.gamma(shape, scale)), lo, hi))


def synth_commit(index, vocab_size=24650, ast_vocab_size=71):
    """-> dict(sou, tar, mark, ast_change, tar_label, sub_token : int arrays; row, col, val : COO)."""
    rng = np.random.default_rng(1_000_003 * (index + 1))
    n_code = _clip_gamma(rng, 4.0, 24.6, 3, 200, offset=2)
    n_sub = _clip_gamma(rng, 3.0, 9.0, 0, 102)
    n_ae = _clip_gamma(rng, 2.2, 14.4, 1, 157)
    n_ast = max(1, int(round(0.8 * n_ae)))
    n_edit = n_ae - n_ast
    n_msg = _clip_gamma(rng, 3.5, 1.66, 1, 20, offset=1)

    sou = np.zeros(N_CODE, np.int64)
    sou[0], sou[n_code - 1] = START, EOS
    sou[1:n_code - 1] = rng.integers(4, vocab_size, n_code - 2)
    mark = np.zeros(N_CODE, np.int64)
    mark[:n_code] = rng.integers(1, 4, n_code)
    mark[0] = mark[n_code - 1] = 2
    sub_token = np.zeros(N_SUB, np.int64)
    sub_token[:n_sub] = rng.integers(4, vocab_size, n_sub)
    ast_change = np.zeros(N_AST, np.int64)
    ast_change[:n_ast] = rng.integers(6, ast_vocab_size, n_ast)
    ast_change[n_ast:n_ast + n_edit] = rng.integers(1, 6, n_edit)

    tar = np.zeros(T_LEN, np.int64)
    tar[0], tar[n_msg + 1] = START, EOS
    tar[1:n_msg + 1] = rng.integers(4, vocab_size, n_msg)
    tar_label = tar.copy()
    kind = rng.random(n_msg)
    for k in range(n_msg):
        if kind[k] < 0.115 and n_code > 2:            # copy from the diff (label = V + position, Dataset.py:202)
            pos = int(rng.integers(1, n_code - 1))
            tar_label[k + 1] = vocab_size + pos
            tar[k + 1] = sou[pos]
        elif kind[k] < 0.25 and n_sub > 0:            # copy from a sub-token (Dataset.py:213)
            pos = int(rng.integers(0, n_sub))
            tar_label[k + 1] = vocab_size + N_CODE + pos
            tar[k + 1] = sub_token[pos]

    pairs = set()

    def link(a, b):
        if a != b:
            pairs.add((a, b)); pairs.add((b, a))

    for j in range(n_code - 1):
        link(j, j + 1)
    for k in range(n_sub):
        for _ in range(1 + (rng.random() < 0.7)):
            link(int(rng.integers(1, max(2, n_code - 1))), N_CODE + k)
    base = N_CODE + N_SUB
    for a in range(n_ast):
        link(base + a, int(rng.integers(1, max(2, n_code - 1))))
        if a > 0 and rng.random() < 0.65:
            link(base + a, base + int(rng.integers(0, a)))
    for c in range(n_edit):
        link(base + n_ast + c, int(rng.integers(1, max(2, n_code - 1))))
        link(base + n_ast + c, base + int(rng.integers(0, n_ast)))
    arr = np.array(sorted(pairs), np.int64).reshape(-1, 2)
    row = np.concatenate((arr[:, 0], np.arange(N_NODES)))
    col = np.concatenate((arr[:, 1], np.arange(N_NODES)))
    deg_r = np.bincount(row, minlength=N_NODES).astype(np.float64)
    deg_c = np.bincount(col, minlength=N_NODES).astype(np.float64)
    val = 1.0 / np.sqrt(deg_r[row]) / np.sqrt(deg_c[col])
    return dict(sou=sou, tar=tar, mark=mark, ast_change=ast_change, tar_label=tar_label, sub_token=sub_token,
                row=row, col=col, val=val)


def synth_batch(first_index, batch_size, vocab_size=24650, ast_vocab_size=71):
    """-> (dict of stacked int64 id arrays, list of (row, col, val) COO triples)."""
    commits = [synth_commit(first_index + i, vocab_size, ast_vocab_size) for i in range(batch_size)]
    ids = {k: np.stack([c[k] for c in commits]) for k in
           ("sou", "tar", "mark", "ast_change", "tar_label", "sub_token")}
    coo = [(c["row"], c["col"], c["val"]) for c in commits]
    return ids, coo


class SynthDataset:
    """`n` synthetic commits in the packed split format of data.TransDataset (`.d` arrays + the length
    attributes), so that data.PackedBatchLoader can serve them exactly like a processed DataSet split."""

    diff_len, sub_token_len, ast_change_len, msg_len = N_CODE, N_SUB, N_AST, T_LEN

    def __init__(self, first_index, n, vocab_size=24650, ast_vocab_size=71):
        commits = [synth_commit(first_index + i, vocab_size, ast_vocab_size) for i in range(n)]
        self.d = {k: np.stack([c[k] for c in commits]).astype(np.int32) for k in
                  ("sou", "tar", "mark", "ast_change", "tar_label", "sub_token")}
        deg, col, val, ptr = [], [], [], [0]
        for c in commits:
            order = np.lexsort((c["col"], c["row"]))                 # CSR order: by row, then by column
            deg.append(np.bincount(c["row"], minlength=N_NODES).astype(np.uint8))
            col.append(c["col"][order].astype(np.int16))
            val.append(c["val"][order])
            ptr.append(ptr[-1] + len(order))
        self.d.update(deg=np.stack(deg), col=np.concatenate(col), val=np.concatenate(val),
                      edge_ptr=np.array(ptr, np.int64))

    def __len__(self):
        return len(self.d["sou"])


def synth_stress_graphs(first_index, batch_size, n_nodes=2048, edges_per_relation=16384, relations=4):
    """BASELINE.json config 5: per graph 4 relations x 16,384 undirected edges drawn uniformly, symmetrised,
    + self loops, degree-normalised.  -> list of (row, col, val)."""
    out = []
    for g in range(batch_size):
        rng = np.random.default_rng(7_000_003 * (first_index + g + 1))
        a = rng.integers(0, n_nodes, relations * edges_per_relation)
        b = rng.integers(0, n_nodes, relations * edges_per_relation)
        keep = a != b
        a, b = a[keep], b[keep]
        key = np.unique(np.concatenate((a * n_nodes + b, b * n_nodes + a)))
        row = np.concatenate((key // n_nodes, np.arange(n_nodes)))
        col = np.concatenate((key % n_nodes, np.arange(n_nodes)))
        deg_r = np.bincount(row, minlength=n_nodes).astype(np.float64)
        deg_c = np.bincount(col, minlength=n_nodes).astype(np.float64)
        out.append((row, col, 1.0 / np.sqrt(deg_r[row]) / np.sqrt(deg_c[col])))
    return out
