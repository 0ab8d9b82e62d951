"""Packed adjacency ("PackedEdges") for the GNN encoder.

The reference hands the model a dense float64 [B, 650, 650] adjacency (Dataset.py:340,
`toarray()`), 0.25 % dense.  The CUDA path consumes a batched CSR instead:

  rowptr int32 [B*N + 1]   cumulative over the batch, rows in (graph b, node i) order
  col    int32 [nnz]       LOCAL source node j in [0, N)
  val    fp32  [nnz]       A[b, i, j]   (the reference casts with edge.float(), gnn_transformer.py:80)

`PackedEdges.from_dense` converts the reference's dense tensor on the device (two passes,
caller-owned buffers); `PackedEdges.from_coo_lists` packs host-side COO lists (what
fira_icse_b200.data emits) without ever materialising the dense matrix.
"""
import numpy as np
import torch

from . import _lib

_EDGE_DTYPE = {torch.float32: _lib.EDGE_F32, torch.float64: _lib.EDGE_F64, torch.bfloat16: _lib.EDGE_BF16}


def _stream():
    return torch.cuda.current_stream().cuda_stream


class PackedEdges:
    """Batched CSR adjacency on one device (+ its transpose when the matrix is not symmetric)."""

    def __init__(self, rowptr, col, val, B, N, symmetric, transpose=None):
        self.rowptr, self.col, self.val = rowptr, col, val
        self.B, self.N = int(B), int(N)
        self.symmetric = bool(symmetric)
        self._t = transpose
        self._rowsum = {}
        self._rows = {}

    @property
    def nnz(self):
        return int(self.col.numel())

    @property
    def device(self):
        return self.rowptr.device

    def t(self):
        """Adjacency of the reversed edges (backward pass).  Dataset.py:346-357 inserts every edge
        in both directions and the value 1/sqrt(deg_r)/sqrt(deg_c) is symmetric, so
        loader-built graphs reuse the forward CSR."""
        if self.symmetric:
            return self
        if self._t is None:
            raise RuntimeError("PackedEdges: transpose not available for a non-symmetric adjacency")
        return self._t

    def rowsum(self, n_code, n_sub, n_ast):
        """sum_j A[b,i,j] per destination row, in the encoder's segment-major row order."""
        key = (n_code, n_sub, n_ast)
        if key not in self._rowsum:
            assert n_code + n_sub + n_ast == self.N
            out = torch.empty(self.B * self.N, dtype=torch.float32, device=self.device)
            _lib.call("fira_csr_rowsum", self.rowptr.data_ptr(), self.val.data_ptr(), self.B, n_code, n_sub,
                      n_ast, out.data_ptr(), _stream())
            self._rowsum[key] = out
        return self._rowsum[key]

    def rows_csr(self, n_code, n_sub, n_ast):
        """The same adjacency as a CSR in BUFFER order for the fused GCN layer kernel (fira_gcn_layer_fwd/bwd):
        rowptr indexed by the segment-major row of the node buffer, col = buffer rows.  -> (rowptr, col, val)."""
        key = (n_code, n_sub, n_ast)
        if self.B == 1:                              # one graph: buffer order == node order, ids are already global
            return self.rowptr, self.col, self.val
        if key not in self._rows:
            assert n_code + n_sub + n_ast == self.N
            R = self.B * self.N
            dev = self.device
            counts = torch.empty(R, dtype=torch.int32, device=dev)
            rowptr = torch.empty(R + 1, dtype=torch.int32, device=dev)
            col = torch.empty(max(self.col.numel(), 1), dtype=torch.int32, device=dev)
            val = torch.empty(max(self.col.numel(), 1), dtype=torch.float32, device=dev)
            _lib.call("fira_csr_to_rows", self.rowptr.data_ptr(), self.col.data_ptr(), self.val.data_ptr(), self.B,
                      n_code, n_sub, n_ast, counts.data_ptr(), rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                      _stream())
            self._rows[key] = (rowptr, col, val)
        return self._rows[key]

    # ------------------------------------------------------------------ constructors
    @staticmethod
    def _csr_from_strided(edge, sb, si, sj, B, N):
        dev = edge.device
        counts = torch.empty(B * N, dtype=torch.int32, device=dev)
        rowptr = torch.empty(B * N + 1, dtype=torch.int32, device=dev)
        et = _EDGE_DTYPE[edge.dtype]
        _lib.call("fira_csr_count_dense", edge.data_ptr(), et, sb, si, sj, B, N, counts.data_ptr(),
                  rowptr.data_ptr(), _stream())
        nnz = int(rowptr[-1].item())            # the one host sync of the dense compatibility path
        col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
        val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)[:nnz]
        _lib.call("fira_csr_fill_dense", edge.data_ptr(), et, sb, si, sj, B, N, rowptr.data_ptr(),
                  col.data_ptr(), val.data_ptr(), _stream())
        return rowptr, col, val

    @classmethod
    def from_dense(cls, edge, assume_symmetric=False):
        """edge: CUDA tensor [B, N, N], float32/float64/bfloat16, any strides."""
        if not edge.is_cuda:
            raise _lib.FiraLibraryError("PackedEdges.from_dense: `edge` must be a CUDA tensor (no CPU fallback)")
        if edge.dtype not in _EDGE_DTYPE:
            edge = edge.float()
        B, N, N2 = edge.shape
        assert N == N2, "adjacency must be square"
        sb, si, sj = edge.stride()
        fwd = cls._csr_from_strided(edge, sb, si, sj, B, N)
        if assume_symmetric:
            return cls(*fwd, B, N, True)
        bwd = cls._csr_from_strided(edge, sb, sj, si, B, N)
        return cls(*fwd, B, N, False, transpose=cls(*bwd, B, N, False))

    @classmethod
    def from_coo_lists(cls, graphs, N, device, symmetric=True, pin=False):
        """graphs: list of (row, col, val) array-likes, one per commit, local node ids.
        Duplicates are summed (scipy `toarray()` semantics, Dataset.py:294,340)."""
        rowptr, col, val = cls.pack_host(graphs, N, pin=pin)
        return cls.from_host(rowptr, col, val, len(graphs), N, device, symmetric)

    @classmethod
    def from_host(cls, rowptr, col, val, B, N, device, symmetric=True):
        """Host (ideally pinned) CSR arrays -> device; three async H2D copies, nothing else."""
        dev = torch.device(device)
        nb = dev.type == "cuda"
        return cls(rowptr.to(dev, non_blocking=nb), col.to(dev, non_blocking=nb), val.to(dev, non_blocking=nb),
                   B, N, symmetric)

    @staticmethod
    def pack_host(graphs, N, pin=False):
        """Collate step of the packed loader: per-commit COO -> batched CSR host tensors."""
        rp, cs, vs = [np.zeros(1, np.int64)], [], []
        base = 0
        for row, col, val in graphs:
            row = np.asarray(row, np.int64); col = np.asarray(col, np.int64); val = np.asarray(val, np.float64)
            key = row * N + col
            uniq, inv = np.unique(key, return_inverse=True)
            if uniq.size != key.size:
                val = np.bincount(inv, weights=val, minlength=uniq.size)
            else:
                val = val[np.argsort(key, kind="stable")]
            r, c = uniq // N, uniq % N
            counts = np.bincount(r, minlength=N)
            rp.append(base + np.cumsum(counts)); base += int(uniq.size)
            cs.append(c.astype(np.int32)); vs.append(val.astype(np.float32))
        rowptr = torch.from_numpy(np.concatenate(rp).astype(np.int32))
        col = torch.from_numpy(np.concatenate(cs) if cs else np.zeros(0, np.int32))
        val = torch.from_numpy(np.concatenate(vs) if vs else np.zeros(0, np.float32))
        if pin:
            rowptr, col, val = rowptr.pin_memory(), col.pin_memory(), val.pin_memory()
        return rowptr, col, val

    def to_dense(self, dtype=torch.float64):
        """Host-side expansion (tests only)."""
        rp = self.rowptr.cpu().numpy(); col = self.col.cpu().numpy(); val = self.val.cpu().numpy()
        out = np.zeros((self.B, self.N, self.N), np.float64)
        for g in range(self.B * self.N):
            b, i = divmod(g, self.N)
            out[b, i, col[rp[g]:rp[g + 1]]] = val[rp[g]:rp[g + 1]]
        return torch.from_numpy(out).to(dtype)
