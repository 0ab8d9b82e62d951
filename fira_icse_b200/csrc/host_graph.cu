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

// ------------------------------------------------------------------------------------------------ per-commit packing
// SURVEY.md 8f rank 4 / Dataset.py:80-94: the reference pads every commit to 210 / 160 / 280 nodes.  Padding nodes are
// isolated (self loop only, Dataset.py:271-275), masked in cross-attention and in the copy softmax, so they never reach
// a real row or the loss.  The packed batch keeps, per commit and segment, only the positions up to the last non-zero
// id ("used" length); node rows are segment-major and RAGGED: [code rows of commit 0, 1, ... | pad | sub-token rows ...
// | pad | AST/edit rows ... | pad], each segment padded to a caller-chosen bucket size with empty rows.
int fira_host_packed_dims(const int* sou, const int* sub_token, const int* ast_change, const unsigned char* deg,
                          const long* index, int batch, int diff_len, int sub_len, int ast_change_len, int* dims) {
  FIRA_CHECK_ARG(sou && sub_token && ast_change && deg && index && dims && batch > 0, FIRA_ERR_ARG,
                 "fira_host_packed_dims: null pointer or empty batch");
  const int n0 = diff_len, n1 = sub_len, n2 = ast_change_len, N = n0 + n1 + n2;
  auto used = [](const int* row, int len) { int u = 0; for (int j = 0; j < len; ++j) if (row[j] != 0) u = j + 1; return u; };
  long rc = 0, rs = 0, ra = 0, nnz = 0;
  int smax = 0, chmax = 0;
  for (int b = 0; b < batch; ++b) {
    const long i = index[b];
    const int uc = used(sou + i * n0, n0), us = used(sub_token + i * n1, n1), ua = used(ast_change + i * n2, n2);
    rc += uc; rs += us; ra += ua;
    smax = std::max(smax, uc + us);
    chmax = std::max(chmax, (uc + 127) / 128 + (us + 127) / 128);        // 128-key chunks the attention kernel needs
    const unsigned char* d = deg + i * N;
    for (int j = 0; j < uc; ++j) nnz += d[j];
    for (int j = 0; j < us; ++j) nnz += d[n0 + j];
    for (int j = 0; j < ua; ++j) nnz += d[n0 + n1 + j];
  }
  dims[0] = (int)rc; dims[1] = (int)rs; dims[2] = (int)ra; dims[3] = smax; dims[4] = (int)nnz; dims[5] = chmax;
  return FIRA_OK;
}

int fira_host_gather_packed(const int* sou, const int* tar, const int* mark, const int* ast_change,
                            const int* tar_label, const int* sub_token, const unsigned char* deg, const short* col,
                            const double* val, const long* edge_ptr, const long* index, int batch, int diff_len,
                            int sub_len, int ast_change_len, int msg_len, int vocab_size, const int* pad_dims,
                            int* o_code, int* o_mark, int* o_pos, int* o_sub, int* o_ast, int* o_off, int* o_ranges,
                            unsigned char* o_mem_mask, int* o_tar, int* o_label, unsigned char* o_tar_mask,
                            int* o_rowptr, int* o_col, float* o_val, long edge_cap, int* nnz_out) {
  FIRA_CHECK_ARG(batch > 0 && diff_len > 0 && msg_len > 0, FIRA_ERR_SHAPE, "fira_host_gather_packed: bad shape");
  FIRA_CHECK_ARG(sou && tar && mark && ast_change && tar_label && sub_token && deg && col && val && edge_ptr && index &&
                     pad_dims && o_code && o_mark && o_pos && o_sub && o_ast && o_off && o_ranges && o_mem_mask &&
                     o_tar && o_label && o_tar_mask && o_rowptr && o_col && o_val && nnz_out,
                 FIRA_ERR_ARG, "fira_host_gather_packed: null pointer");
  const int n0 = diff_len, n1 = sub_len, n2 = ast_change_len, N = n0 + n1 + n2, V = vocab_size;
  const int Rc = pad_dims[0], Rs = pad_dims[1], Ra = pad_dims[2], S = pad_dims[3];
  auto used = [](const int* row, int len) { int u = 0; for (int j = 0; j < len; ++j) if (row[j] != 0) u = j + 1; return u; };
  int* off_c = o_off; int* off_s = o_off + (batch + 1); int* off_a = o_off + 2 * (batch + 1);
  off_c[0] = off_s[0] = off_a[0] = 0;
  for (int b = 0; b < batch; ++b) {
    const long i = index[b];
    off_c[b + 1] = off_c[b] + used(sou + i * n0, n0);
    off_s[b + 1] = off_s[b] + used(sub_token + i * n1, n1);
    off_a[b + 1] = off_a[b] + used(ast_change + i * n2, n2);
  }
  FIRA_CHECK_ARG(off_c[batch] <= Rc && off_s[batch] <= Rs && off_a[batch] <= Ra, FIRA_ERR_SHAPE,
                 "fira_host_gather_packed: batch needs (%d, %d, %d) rows, buffers hold (%d, %d, %d)", off_c[batch],
                 off_s[batch], off_a[batch], Rc, Rs, Ra);
  // ---- node ids (int32), padding rows of each segment = id 0
  memset(o_code, 0, sizeof(int) * Rc); memset(o_mark, 0, sizeof(int) * Rc); memset(o_pos, 0, sizeof(int) * Rc);
  memset(o_sub, 0, sizeof(int) * Rs); memset(o_ast, 0, sizeof(int) * Ra);
  memset(o_mem_mask, 0, (size_t)batch * S);
  for (int b = 0; b < batch; ++b) {
    const long i = index[b];
    const int uc = off_c[b + 1] - off_c[b], us = off_s[b + 1] - off_s[b], ua = off_a[b + 1] - off_a[b];
    FIRA_CHECK_ARG(uc + us <= S, FIRA_ERR_SHAPE, "fira_host_gather_packed: commit %ld has %d memory rows, S = %d", i,
                   uc + us, S);
    for (int j = 0; j < uc; ++j) {
      o_code[off_c[b] + j] = sou[i * n0 + j];
      o_mark[off_c[b] + j] = mark[i * n0 + j];
      o_pos[off_c[b] + j] = j;
      o_mem_mask[(long)b * S + j] = sou[i * n0 + j] != 0;          // an interior zero id stays masked (Model.py:42)
    }
    for (int j = 0; j < us; ++j) {
      o_sub[off_s[b] + j] = sub_token[i * n1 + j];
      o_mem_mask[(long)b * S + uc + j] = sub_token[i * n1 + j] != 0;
    }
    for (int j = 0; j < ua; ++j) o_ast[off_a[b] + j] = ast_change[i * n2 + j];
    o_ranges[4 * b + 0] = off_c[b]; o_ranges[4 * b + 1] = uc;
    o_ranges[4 * b + 2] = Rc + off_s[b]; o_ranges[4 * b + 3] = us;
    // decoder input, shifted labels (Model.py:71-79) with copy labels renumbered to the commit's own memory rows:
    // code position s -> V + s, sub-token k -> V + uc + k; a label on a padded source position (p = 0 in the
    // reference: masked, Model.py:61) -> V + S, which head_fwd_kernel treats as "beyond the source" (p = 0 as well)
    for (int t = 0; t < msg_len; ++t) {
      const int tk = tar[i * msg_len + t];
      o_tar[(long)b * msg_len + t] = tk;
      o_tar_mask[(long)b * msg_len + t] = tk != 0;
      long l = t + 1 < msg_len ? tar_label[i * msg_len + t + 1] : 0;
      if (l >= V) {
        const long s = l - V;
        if (s < n0) l = s < uc ? V + s : (long)V + S;
        else { const long k = s - n0; l = k < us ? V + uc + k : (long)V + S; }
      }
      o_label[(long)b * msg_len + t] = (int)l;
    }
  }
  // ---- buffer-order CSR with GLOBAL column ids
  long nnz = 0;
  o_rowptr[0] = 0;
  const int seg_lo[3] = {0, n0, n0 + n1};
  const int seg_base[3] = {0, Rc, Rc + Rs};
  const int seg_rows[3] = {Rc, Rs, Ra};
  const int* seg_off[3] = {off_c, off_s, off_a};
  for (int sgm = 0; sgm < 3; ++sgm) {
    for (int b = 0; b < batch; ++b) {
      const long i = index[b];
      const unsigned char* d = deg + i * N;
      long e = edge_ptr[i];
      for (int r = 0; r < seg_lo[sgm]; ++r) e += d[r];
      const int u = seg_off[sgm][b + 1] - seg_off[sgm][b];
      const int ucs[3] = {off_c[b + 1] - off_c[b], off_s[b + 1] - off_s[b], off_a[b + 1] - off_a[b]};
      for (int j = 0; j < u; ++j) {
        const int dr = d[seg_lo[sgm] + j];
        FIRA_CHECK_ARG(nnz + dr <= edge_cap, FIRA_ERR_SHAPE, "fira_host_gather_packed: more than %ld edges", edge_cap);
        for (int k = 0; k < dr; ++k) {
          const int c = col[e + k];
          const int cs = c < n0 ? 0 : (c < n0 + n1 ? 1 : 2);
          const int cj = c - seg_lo[cs];
          FIRA_CHECK_ARG(cj < ucs[cs], FIRA_ERR_ARG,
                         "fira_host_gather_packed: commit %ld node %d has a neighbour (%d) inside the dropped padding", i,
                         seg_lo[sgm] + j, c);
          o_col[nnz + k] = seg_base[cs] + seg_off[cs][b] + cj;
          o_val[nnz + k] = (float)val[e + k];
        }
        nnz += dr;
        e += dr;
        o_rowptr[seg_base[sgm] + seg_off[sgm][b] + j + 1] = (int)nnz;
      }
    }
    for (int r = seg_off[sgm][batch]; r < seg_rows[sgm]; ++r) o_rowptr[seg_base[sgm] + r + 1] = (int)nnz;   // empty pad rows
  }
  *nnz_out = (int)nnz;
  return FIRA_OK;
}

}  // extern "C"
