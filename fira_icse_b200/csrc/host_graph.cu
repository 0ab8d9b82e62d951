// Host-side batch preparation (no GPU work, no stream): the commit-graph adjacency assembly of
// Dataset.py:220-294 + process_edge (Dataset.py:346-357), and the loader's gather / collate / padding
// trim (Dataset.py:336-343, run_model.py:387) as one pass that writes straight into caller-owned
// (pinned) staging buffers.  Plain C++; compiled into libfira_b200.so next to the kernels.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "fira_b200.h"

namespace {

// one undirected relation: pairs (a, b) -> node ids (a + off_a, b + off_b); `limit_b` drops pairs whose
// second node id is not below it (the reference drops edges to code tokens beyond the padded length).
inline void add_pairs(std::vector<int>& und, const int* pairs, int n, int off_a, int off_b, int limit_b) {
  for (int e = 0; e < n; ++e) {
    const int a = pairs[2 * e] + off_a, b = pairs[2 * e + 1] + off_b;
    if (limit_b >= 0 && b >= limit_b) continue;
    und.push_back(a);
    und.push_back(b);
  }
}

}  // namespace

extern "C" {

int fira_host_build_adjacency(const int* change_code, int n_change_code, const int* change_ast, int n_change_ast,
                              const int* ast_code, int n_ast_code, const int* ast_ast, int n_ast_ast,
                              const int* code_sub, int n_code_sub, int n_diff, int n_ast, int diff_len, int sub_len,
                              int ast_change_len, int* deg, int* col, double* val, int cap, int* nnz_out) {
  FIRA_CHECK_ARG(diff_len > 0 && sub_len >= 0 && ast_change_len >= 0 && n_diff >= 0 && n_ast >= 0, FIRA_ERR_SHAPE,
                 "fira_host_build_adjacency: bad lengths");
  FIRA_CHECK_ARG(deg && col && val && nnz_out, FIRA_ERR_ARG, "fira_host_build_adjacency: null output");
  const int n_nodes = diff_len + sub_len + ast_change_len;
  const int a0 = diff_len + sub_len;                       // first AST node; edit nodes follow the n_ast AST nodes
  std::vector<int> und;
  und.reserve(2 * (size_t)(n_change_code + n_change_ast + n_ast_code + n_ast_ast + n_code_sub + n_diff + 1));
  add_pairs(und, change_code, n_change_code, a0 + n_ast, 1, diff_len);      // edit  - code  (code j -> j+1)
  add_pairs(und, change_ast, n_change_ast, a0 + n_ast, a0, -1);            // edit  - AST
  add_pairs(und, ast_code, n_ast_code, a0, 1, diff_len);                    // AST   - code
  add_pairs(und, ast_ast, n_ast_ast, a0, a0, -1);                           // AST   - AST
  add_pairs(und, code_sub, n_code_sub, 1, diff_len, -1);                    // code  - sub-token
  for (int j = 0; j <= n_diff; ++j) {                                       // <start> t1 ... tn <eos> chain
    und.push_back(j);
    und.push_back(j + 1);
  }
  // ordered pairs keyed r * n + c: both directions + the self loop of every node, de-duplicated
  std::vector<long> keys;
  keys.reserve(und.size() + n_nodes);
  for (size_t e = 0; e < und.size(); e += 2) {
    const long a = und[e], b = und[e + 1];
    FIRA_CHECK_ARG(a >= 0 && b >= 0 && a < n_nodes && b < n_nodes, FIRA_ERR_SHAPE,
                   "fira_host_build_adjacency: node id (%ld, %ld) outside the %d-node graph", a, b, n_nodes);
    FIRA_CHECK_ARG(a != b, FIRA_ERR_ARG, "fira_host_build_adjacency: self edge %ld in the input relations", a);
    keys.push_back(a * n_nodes + b);
    keys.push_back(b * n_nodes + a);
  }
  for (long i = 0; i < n_nodes; ++i) keys.push_back(i * n_nodes + i);
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  const int nnz = (int)keys.size();
  *nnz_out = nnz;
  FIRA_CHECK_ARG(nnz <= cap, FIRA_ERR_SHAPE, "fira_host_build_adjacency: %d entries, capacity %d", nnz, cap);
  // the matrix is symmetric: row degree == column degree
  memset(deg, 0, sizeof(int) * n_nodes);
  for (int e = 0; e < nnz; ++e) deg[keys[e] / n_nodes]++;
  for (int e = 0; e < nnz; ++e) {
    const int r = (int)(keys[e] / n_nodes), c = (int)(keys[e] % n_nodes);
    col[e] = c;
    val[e] = 1.0 / sqrt((double)deg[r]) / sqrt((double)deg[c]);            // Dataset.py:277-291, float64
  }
  return FIRA_OK;
}

int fira_host_batch_dims(const int* sou, const int* sub_token, const int* ast_change, const long* index, int batch,
                         int diff_len, int sub_len, int ast_change_len, int mult_code, int mult_sub, int mult_ast,
                         int* dims) {
  FIRA_CHECK_ARG(sou && sub_token && ast_change && index && dims && batch > 0, FIRA_ERR_ARG,
                 "fira_host_batch_dims: null pointer or empty batch");
  auto used = [&](const int* base, int len, int mult) {
    if (mult <= 0) return len;                            // no trimming of this segment
    int m = 0;
    for (int b = 0; b < batch; ++b) {
      const int* row = base + index[b] * len;
      int u = 0;
      for (int j = 0; j < len; ++j)
        if (row[j] != 0) u = j + 1;                       // position after the last non-padding id
      m = std::max(m, u);
    }
    return std::min(len, std::max(mult, (m + mult - 1) / mult * mult));
  };
  dims[0] = used(sou, diff_len, mult_code);
  dims[1] = used(sub_token, sub_len, mult_sub);
  dims[2] = used(ast_change, ast_change_len, mult_ast);
  return FIRA_OK;
}

int fira_host_gather_batch(const int* sou, const int* tar, const int* mark, const int* ast_change,
                           const int* tar_label, const int* sub_token, const unsigned char* deg, const short* col,
                           const double* val, const long* edge_ptr, const long* index, int batch, int diff_len,
                           int sub_len, int ast_change_len, int msg_len, int vocab_size, const int* dims,
                           long* o_sou, long* o_tar, long* o_mark, long* o_ast_change, long* o_tar_label,
                           long* o_sub_token, int* o_rowptr, int* o_col, float* o_val, long edge_cap, int* nnz_out) {
  FIRA_CHECK_ARG(batch > 0 && diff_len > 0 && msg_len > 0, FIRA_ERR_SHAPE, "fira_host_gather_batch: bad shape");
  FIRA_CHECK_ARG(sou && tar && mark && ast_change && tar_label && sub_token && deg && col && val && edge_ptr && index &&
                     o_sou && o_tar && o_mark && o_ast_change && o_tar_label && o_sub_token && o_rowptr && o_col &&
                     o_val && dims && nnz_out,
                 FIRA_ERR_ARG, "fira_host_gather_batch: null pointer");
  const int n0 = diff_len, n1 = sub_len, n2 = ast_change_len, N = n0 + n1 + n2;
  const int c0 = dims[0], c1 = dims[1], c2 = dims[2];
  FIRA_CHECK_ARG(c0 > 0 && c0 <= n0 && c1 >= 0 && c1 <= n1 && c2 >= 0 && c2 <= n2, FIRA_ERR_SHAPE,
                 "fira_host_gather_batch: segment lengths (%d, %d, %d) outside (%d, %d, %d)", c0, c1, c2, n0, n1, n2);
  // ids cut away must be padding (a real token there means dims came from a different batch)
  for (int b = 0; b < batch; ++b) {
    const long i = index[b];
    bool clean = true;
    for (int j = c0; j < n0; ++j) clean &= sou[i * n0 + j] == 0;
    for (int j = c1; j < n1; ++j) clean &= sub_token[i * n1 + j] == 0;
    for (int j = c2; j < n2; ++j) clean &= ast_change[i * n2 + j] == 0;
    FIRA_CHECK_ARG(clean, FIRA_ERR_ARG, "fira_host_gather_batch: commit %ld has real tokens beyond (%d, %d, %d)", i, c0,
                   c1, c2);
  }
  const int Nt = c0 + c1 + c2;
  std::vector<int> remap(N, -1);
  for (int j = 0; j < c0; ++j) remap[j] = j;
  for (int j = 0; j < c1; ++j) remap[n0 + j] = c0 + j;
  for (int j = 0; j < c2; ++j) remap[n0 + n1 + j] = c0 + c1 + j;
  // ---- id tensors (int64 like the reference's collate), labels renumbered for the shorter code segment
  for (int b = 0; b < batch; ++b) {
    const long i = index[b];
    for (int j = 0; j < c0; ++j) {
      o_sou[(long)b * c0 + j] = sou[i * n0 + j];
      o_mark[(long)b * c0 + j] = mark[i * n0 + j];
    }
    for (int j = 0; j < c1; ++j) o_sub_token[(long)b * c1 + j] = sub_token[i * n1 + j];
    for (int j = 0; j < c2; ++j) o_ast_change[(long)b * c2 + j] = ast_change[i * n2 + j];
    for (int j = 0; j < msg_len; ++j) {
      o_tar[(long)b * msg_len + j] = tar[i * msg_len + j];
      long l = tar_label[i * msg_len + j];
      if (l >= vocab_size + n0) l -= (n0 - c0);          // sub-token copy labels sit behind the code segment
      // a label that lands beyond c0 + c1 pointed at a PADDED source position of the untrimmed batch (masked there:
      // p = 0 -> clamp floor, Model.py:61,69); head_fwd_kernel gives such labels the same p = 0 without reading them
      o_tar_label[(long)b * msg_len + j] = l;
    }
  }
  // ---- batch CSR over the kept rows
  long nnz = 0;
  o_rowptr[0] = 0;
  for (int b = 0; b < batch; ++b) {
    const long i = index[b];
    const unsigned char* d = deg + i * N;
    long e = edge_ptr[i];
    for (int r = 0; r < N; ++r) {
      const int dr = d[r];
      if (remap[r] >= 0) {
        FIRA_CHECK_ARG(nnz + dr <= edge_cap, FIRA_ERR_SHAPE, "fira_host_gather_batch: more than %ld edges", edge_cap);
        for (int k = 0; k < dr; ++k) {
          const int c = remap[col[e + k]];
          FIRA_CHECK_ARG(c >= 0, FIRA_ERR_ARG,
                         "fira_host_gather_batch: commit %ld node %d has a neighbour inside the trimmed padding", i, r);
          o_col[nnz + k] = c;
          o_val[nnz + k] = (float)val[e + k];
        }
        nnz += dr;
        o_rowptr[(long)b * Nt + remap[r] + 1] = (int)nnz;
      }
      e += dr;
    }
    FIRA_CHECK_ARG(e == edge_ptr[i + 1], FIRA_ERR_ARG, "fira_host_gather_batch: degree table and edge_ptr disagree");
  }
  *nnz_out = (int)nnz;
  return FIRA_OK;
}

}  // extern "C"
