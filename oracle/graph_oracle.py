"""CPU oracle for the commit -> padded-id / adjacency builder -- TEST INFRASTRUCTURE ONLY.

Pure-Python restatement of the per-commit part of the reference's
``TransDataset.process_data`` (/root/reference/Dataset.py:96-294) and ``process_edge``
(Dataset.py:346-357).  It is the checker for fira_icse_b200.graph (the product packer),
and is itself pinned to the reference's own output by tests/test_oracle_golden.py via
tests/golden/batch_first128.npz and by tests/test_data.py via tests/golden/batch_edge.npz
(DataSet extremes and crafted commits that reach the truncation branches).

Node numbering (Dataset.py:220-266): code token j -> j+1 (0 is <start>), sub-token k ->
210+k, AST node a -> 370+a, edit node c -> 370+len(ast)+c.  Every relation is inserted
in both directions, de-duplicated on the ordered pair; every one of the 650 nodes gets
a self loop; value = 1/sqrt(deg_row)/sqrt(deg_col) in float64 (Dataset.py:277-291).
"""
import math

LEMMA = {"added": "add", "fixed": "fix", "removed": "remove",
         "adding": "add", "fixing": "fix", "removing": "remove"}   # Dataset.py:15


def _ids(tokens, vocab, upper):
    """Dataset.py:70-79: lower-case unless white-listed, unknown -> <unkm>."""
    out = []
    for t in tokens:
        if t not in upper:
            t = t.lower()
        out.append(vocab[t] if t in vocab else vocab["<unkm>"])
    return out


def _pad(seq, n, pad=0):
    return (seq + [pad] * (n - len(seq)))[:n]


def build_commit(raw, i, vocab, ast_vocab, upper, diff_len=210, msg_len=30, att_len=25,
                 ast_change_len=280, sub_len=160):
    """Return dict(sou, tar, attr, mark, ast_change, tar_label, sub_token, row, col, val)
    for commit i of the raw lists (keys as the DataSet/*.json file stems)."""
    var_map = raw["variable"][i]
    diff = list(raw["difftoken"][i])
    msg = list(raw["msg"][i])
    atts = raw["diffatt"][i]
    # anonymise + lower-case (Dataset.py:122-135)
    for j, t in enumerate(diff):
        t = var_map.get(t, t)
        diff[j] = t if t in upper else t.lower()
    for j, t in enumerate(msg):
        t = var_map.get(t, t)
        t = t if t in upper else t.lower()
        msg[j] = LEMMA.get(t, t)

    sou = _pad([vocab["<start>"]] + _ids(diff, vocab, upper) + [vocab["<eos>"]], diff_len)
    label = _ids(msg, vocab, upper)
    tar = _pad([vocab["<start>"]] + list(label) + [vocab["<eos>"]], msg_len)
    attr = [[]] + [_ids(a, vocab, upper) for a in atts] + [[]]
    attr = [_pad(a, att_len) for a in attr]
    attr = (attr + [[0] * att_len] * (diff_len - len(attr)))[:diff_len]
    mark = _pad([2] + list(raw["diffmark"][i]) + [2], diff_len)
    n_ast = len(raw["ast"][i])
    ast_change = _pad(_ids(list(raw["ast"][i]) + list(raw["change"][i]), ast_vocab, upper), ast_change_len)

    # sub-token nodes: one set per distinct identifier, first occurrence wins (Dataset.py:173-192)
    sub_tokens, first_seen, sub_edges = [], {}, []
    for j, att in enumerate(atts):
        if not att:
            continue
        tok = diff[j]
        if tok not in first_seen:
            first_seen[tok] = list(range(len(sub_tokens), len(sub_tokens) + len(att)))
            sub_tokens += list(att)
        for k in first_seen[tok]:
            sub_edges.append((j, k))
    sub_token = _pad(_ids(sub_tokens, vocab, upper), sub_len)

    # dual-copy labels: diff position preferred, else sub-token position (Dataset.py:199-213)
    V = len(vocab)
    for k, w in enumerate(msg):
        if w in diff:
            label[k] = diff.index(w) + V + 1
    for k, w in enumerate(msg):
        if w in sub_tokens and label[k] < V:
            label[k] = sub_tokens.index(w) + V + diff_len
    tar_label = _pad([vocab["<start>"]] + label + [vocab["<eos>"]], msg_len)

    # adjacency (Dataset.py:220-294)
    seen, row, col = set(), [], []

    def link(a, b):
        for p, q in ((a, b), (b, a)):
            if (p, q) not in seen:
                seen.add((p, q)); row.append(p); col.append(q)

    base_ast = diff_len + sub_len
    for c, j in raw["edge_change_code"][i]:
        if j + 1 < diff_len:
            link(c + base_ast + n_ast, j + 1)
    for c, a in raw["edge_change_ast"][i]:
        link(c + base_ast + n_ast, a + base_ast)
    for a, j in raw["edge_ast_code"][i]:
        if j + 1 < diff_len:
            link(a + base_ast, j + 1)
    for a, b in raw["edge_ast"][i]:
        link(a + base_ast, b + base_ast)
    for j, k in sub_edges:
        link(j + 1, k + diff_len)
    for j in range(len(diff) + 1):
        link(j, j + 1)
    n = diff_len + sub_len + ast_change_len
    for v in range(n):
        assert (v, v) not in seen
        row.append(v); col.append(v)
    deg_r, deg_c = [0] * n, [0] * n
    for r in row:
        deg_r[r] += 1
    for c in col:
        deg_c[c] += 1
    val = [1 / math.sqrt(deg_r[r]) / math.sqrt(deg_c[c]) for r, c in zip(row, col)]
    return dict(sou=sou, tar=tar, attr=attr, mark=mark, ast_change=ast_change,
                tar_label=tar_label, sub_token=sub_token, row=row, col=col, val=val)
