"""Per-commit packed batches (fira_host_packed_dims / fira_host_gather_packed, fira_icse_b200/packed.py) against the
padded batch the reference's collate produces (golden commits processed by the reference's own process_data):
every real node row, every adjacency entry, every label is recovered from the packed form; only padding is gone."""
import numpy as np
import pytest
import torch

from fira_testlib import load_batch_golden

V = 24650
N0, N1, N2, T = 210, 160, 280, 30


class GoldenSplit:
    """the 128 golden commits in the packed split format (what data.TransDataset.d holds)"""

    def __init__(self):
        g = load_batch_golden()
        ptr = g["edge_ptr"]
        n = len(g["sou"])
        deg = np.zeros((n, 650), np.uint8)
        cols, vals, eptr = [], [], [0]
        for i in range(n):
            r, c, v = g["edge_row"][ptr[i]:ptr[i + 1]], g["edge_col"][ptr[i]:ptr[i + 1]], g["edge_val"][ptr[i]:ptr[i + 1]]
            order = np.lexsort((c, r))
            r, c, v = r[order], c[order], v[order]
            deg[i] = np.bincount(r, minlength=650)
            cols.append(c.astype(np.int16)); vals.append(v.astype(np.float64)); eptr.append(eptr[-1] + len(c))
        self.d = {k: g[k].astype(np.int32) for k in ("sou", "tar", "mark", "ast_change", "tar_label", "sub_token")}
        self.d.update(deg=deg, col=np.concatenate(cols), val=np.concatenate(vals), edge_ptr=np.array(eptr, np.int64))
        self.diff_len, self.sub_token_len, self.ast_change_len, self.msg_len = N0, N1, N2, T
        self.g = g

    def __len__(self):
        return len(self.d["sou"])


def _used(row):
    nz = np.nonzero(row)[0]
    return int(nz.max()) + 1 if nz.size else 0


@pytest.mark.parametrize("index", [list(range(0, 12)), [5], [100, 3, 77, 3, 127, 64, 9]])
def test_packed_batch_recovers_the_padded_batch(index):
    from fira_icse_b200.packed import PackedTables, pack_from_dataset
    ds = GoldenSplit()
    tables = PackedTables(ds)
    pb = pack_from_dataset(tables, np.array(index), V)
    B = len(index)
    d = ds.d
    uc = [_used(d["sou"][i]) for i in index]
    us = [_used(d["sub_token"][i]) for i in index]
    ua = [_used(d["ast_change"][i]) for i in index]
    need = tables.dims(np.array(index))
    assert need[:3] == (sum(uc), sum(us), sum(ua)) and need[3] == max(a + b for a, b in zip(uc, us))
    assert pb.Rc % 1024 == 0 and pb.Rs % 512 == 0 and pb.Ra % 512 == 0 and pb.S % 64 == 0
    assert pb.Rc >= sum(uc) and pb.S >= need[3] and pb.nnz == need[4]
    off = pb.off.numpy()
    assert list(off[0]) == list(np.concatenate(([0], np.cumsum(uc)))) and list(off[1]) == list(np.concatenate(([0], np.cumsum(us))))
    code, mark, pos, sub, ast = (getattr(pb, k).numpy() for k in ("code", "mark", "pos", "sub", "ast"))
    rowptr, col, val = pb.rowptr.numpy(), pb.col.numpy(), pb.val.numpy()
    assert rowptr[0] == 0 and rowptr[-1] == pb.nnz and (np.diff(rowptr) >= 0).all()
    ranges = pb.ranges.numpy()
    for b, i in enumerate(index):
        assert np.array_equal(code[off[0][b]:off[0][b + 1]], d["sou"][i][:uc[b]])
        assert np.array_equal(mark[off[0][b]:off[0][b + 1]], d["mark"][i][:uc[b]])
        assert np.array_equal(pos[off[0][b]:off[0][b + 1]], np.arange(uc[b]))
        assert np.array_equal(sub[off[1][b]:off[1][b + 1]], d["sub_token"][i][:us[b]])
        assert np.array_equal(ast[off[2][b]:off[2][b + 1]], d["ast_change"][i][:ua[b]])
        assert list(ranges[b]) == [off[0][b], uc[b], pb.Rc + off[1][b], us[b]]
        mm = pb.mem_mask.numpy()[b]
        expect = np.concatenate((d["sou"][i][:uc[b]] != 0, d["sub_token"][i][:us[b]] != 0))
        assert np.array_equal(mm[:uc[b] + us[b]], expect.astype(np.uint8)) and not mm[uc[b] + us[b]:].any()
        assert np.array_equal(pb.tar.numpy()[b], d["tar"][i]) and np.array_equal(pb.tar_mask.numpy()[b], d["tar"][i] != 0)
        # shifted labels, copy labels renumbered to the commit's own memory rows
        lab = np.concatenate((d["tar_label"][i][1:], [0]))
        exp = lab.copy()
        for t, l in enumerate(lab):
            if l >= V:
                s = l - V
                exp[t] = V + s if s < N0 else V + uc[b] + (s - N0)
                assert (s < uc[b]) if s < N0 else (s - N0 < us[b])          # golden labels point at real positions
        assert np.array_equal(pb.label.numpy()[b], exp)
        # adjacency: dense [650,650] of the commit restricted to its real nodes == the packed rows mapped back
        dense = np.zeros((650, 650))
        lo, hi = d["edge_ptr"][i], d["edge_ptr"][i + 1]
        dense[np.repeat(np.arange(650), d["deg"][i]), d["col"][lo:hi]] = d["val"][lo:hi]
        node_of_row = {}
        for j in range(uc[b]):
            node_of_row[off[0][b] + j] = j
        for j in range(us[b]):
            node_of_row[pb.Rc + off[1][b] + j] = N0 + j
        for j in range(ua[b]):
            node_of_row[pb.Rc + pb.Rs + off[2][b] + j] = N0 + N1 + j
        rebuilt = np.zeros((650, 650), np.float32)
        for r, node in node_of_row.items():
            for e in range(rowptr[r], rowptr[r + 1]):
                rebuilt[node, node_of_row[col[e]]] = val[e]
        real = np.zeros(650, bool)
        real[list(node_of_row.values())] = True
        assert np.array_equal(rebuilt[np.ix_(real, real)], dense[np.ix_(real, real)].astype(np.float32))
        # what was dropped is pure padding: isolated self loops
        assert np.array_equal(dense[~real][:, ~real], np.eye((~real).sum())) and not dense[np.ix_(real, ~real)].any()
    # padding rows of every segment are empty
    for lo, hi in ((off[0][B], pb.Rc), (pb.Rc + off[1][B], pb.Rc + pb.Rs), (pb.Rc + pb.Rs + off[2][B], pb.rows)):
        assert (np.diff(rowptr[lo:hi + 1]) == 0).all()
    assert not code[off[0][B]:].any() and not sub[off[1][B]:].any() and not ast[off[2][B]:].any()


def test_packed_row_counts_on_the_goldens():
    """the point of packing: ~4x fewer node rows than 650 per commit, ~2.4x fewer than batch-level trimming"""
    from fira_icse_b200.packed import PackedTables
    ds = GoldenSplit()
    tables = PackedTables(ds)
    need = tables.dims(np.arange(64))
    rows = sum(need[:3])
    assert rows < 64 * 650 / 3
    print("64 golden commits: packed rows", rows, "vs padded", 64 * 650, "memory rows max per commit", need[3])


def test_packed_gather_rejects_too_small_buffers():
    from fira_icse_b200 import FiraLibraryError
    from fira_icse_b200.packed import PackedTables, pack_from_dataset
    tables = PackedTables(GoldenSplit())
    with pytest.raises(FiraLibraryError, match="rows"):
        pack_from_dataset(tables, np.arange(8), V, pad_dims=(64, 512, 512, 384))
