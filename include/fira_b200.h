/* libfira_b200 -- C ABI of the B200-native FIRA hot path.
 *
 * The reference (DJjjjhao/FIRA-ICSE) has no FFI of its own: its hot path is PyTorch library
 * calls issued from Model.py / gnn_transformer.py / combination_layer.py.  The drop-in boundary
 * is therefore the nn.Module surface (kept by fira_icse_b200/), and THIS header is the thin
 * C ABI those modules call instead of torch ops.  Each entry point names the reference
 * statement(s) it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory owned by the caller (PyTorch caching
 *     allocator); the library never allocates, frees or retains a pointer;
 *   - tensors are contiguous row-major unless a leading dimension (ld*) is given, 16-byte aligned;
 *   - ids / indices are int32; masks are uint8 (1 = keep);
 *   - `dtype` selects the ACTIVATION storage type: FIRA_F32 (parity mode) or FIRA_BF16 (throughput
 *     mode); parameters, statistics and gradients of parameters are always fp32;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous, no implicit sync;
 *   - return 0 on success, a FIRA_ERR_* code otherwise; fira_last_error_string() describes the last
 *     failure on the calling thread; nothing throws or aborts across the boundary;
 *   - re-entrant, no hidden global state besides the per-thread error string;
 *   - dropout masks are a pure function of (seed [+ *seed_ctr], stream_id, element index): backward
 *     recomputes them; seed_ctr (device uint64, may be NULL) lets a captured CUDA graph draw fresh masks
 *     on every replay by bumping the counter on the device.
 */
#ifndef FIRA_B200_H_
#define FIRA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FIRA_F32 0
#define FIRA_BF16 1
#define FIRA_EDGE_F32 0
#define FIRA_EDGE_BF16 1
#define FIRA_EDGE_F64 2

int fira_version(void);                    /* ABI version, bumped on any signature change */
const char* fira_last_error_string(void);
int fira_built_arch(void);                 /* 100 when compiled for sm_100a */
/* Launch mode of every kernel of the library (the one process-wide switch, atomic): on = programmatic dependent
 * launch -- a kernel's CTAs are scheduled while the previous kernel of the stream drains and block in
 * griddepcontrol.wait before their first global-memory access (results are identical).
 * Default: on (FIRA_PDL=0 in the environment turns it off).  Measured on the captured training step (profiles/):
 * 3.23 -> 3.05 ms once the side work runs on several streams; with ONE side stream it was 3% slower. */
int fira_set_pdl(int on);
int fira_get_pdl(void);

/* ---- generic fp32 Linear pieces (every nn.Linear of the path; e.g. gnn_transformer.py:78,82,
 *      141-143,158,171-173,200-204; Model.py:16-19,54).
 *      C[M,N] = A(MxK) * B(KxN) + bias[n] + rs[m]*rc[n], optional relu.
 *      A(m,k) = a_kcontig ? A[m*lda+k] : A[k*lda+m];  B(k,n) = b_kcontig ? B[n*ldb+k] : B[k*ldb+n].
 *      accumulate: C += result.  splits>1: K is split across CTAs, partials are atomically added
 *      (the library zero-fills C first unless accumulate). */
int fira_gemm_f32(const float* A, long lda, int a_kcontig, const float* B, long ldb, int b_kcontig, float* C,
                  long ldc, int M, int N, int K, const float* bias, const float* rs, const float* rc, int relu,
                  int accumulate, int splits, void* stream);

/* ---- bf16 tensor-core Linear (throughput mode): tcgen05.mma with TMEM accumulators, TMA-staged
 *      operands.  Same contraction as fira_gemm_f32 on bf16 operands (fp32 accumulate):
 *      A(m,k) = a_kmajor ? A[m*lda+k] : A[k*lda+m];  B(k,n) = b_kmajor ? B[n*ldb+k] : B[k*ldb+n];
 *      lda/ldb multiples of 8; C fp32 or bf16 (c_is_bf16); accumulate: C += result;
 *      splits>1 = split-K with fp32 atomics (C zero-filled first unless accumulate). */
int fira_gemm_bf16_tc(const void* A, long lda, int a_kmajor, const void* B, long ldb, int b_kmajor, void* C,
                      long ldc, int c_is_bf16, int M, int N, int K, const float* bias, const float* rs,
                      const float* rc, int relu, int accumulate, int splits, void* stream);

/* Weight-gradient form with the bias gradient folded in (every `dW = dY^T X`, `db = colsum(dY)` pair of the backward
 * pass, e.g. gnn_transformer.py:141-143,158,171-173 under autograd): A = dY read MN-major (A(m,k) = A[k*lda+m]),
 * C[M,N] = A B as above, and d_bias[m] += sum_k A(m,k), accumulated atomically into a zero-filled buffer from the A
 * tiles while they sit in shared memory -- no separate column-sum pass over dY. */
int fira_gemm_bf16_tc_dbias(const void* A, long lda, const void* B, long ldb, int b_kmajor, void* C, long ldc,
                            int c_is_bf16, int M, int N, int K, int accumulate, int splits, float* d_bias, void* stream);

/* Linear + dropout + residual + LayerNorm in one launch (bf16 throughput mode): the block that closes every sub-layer
 * (gnn_transformer.py:83 GCN, :158-161 Attention, :173-174 FeedForward, :204-205 Combination).
 *   z[m,:] = x[m,:] W^T + bias (+ rs[m] * rc[:])     W: bf16 [256, K] row-major, x: bf16 [rows, K] (leading dim ldx)
 *   out    = LN(dropout_p(z) + resid) * gamma + beta   rows < split -> outA[r], the others -> outB[r] (global row index;
 *            outB may be NULL: every row goes to outA; outB must hold all `rows` rows -- the 32-row slab that straddles
 *            `split` is stored whole, so up to 31 rows below `split` of outB are written too);  z (bf16 [rows,256]), mean / rstd (fp32 [rows]) are kept for
 *            fira_ln_residual_bwd.  Same dropout masks as fira_ln_residual_fwd with the same (seed, stream_id). */
int fira_gemm_ln_fwd(const void* x, long ldx, const void* w, const float* bias, const float* rs, const float* rc,
                     const void* resid, const float* gamma, const float* beta, void* z, void* outA, void* outB, long split,
                     float* mean, float* rstd, long rows, int K, float p_drop, uint64_t seed, const uint64_t* seed_ctr,
                     uint32_t stream_id, void* stream);

/* Input gradient through a relu (gnn_transformer.py:172 under autograd): dx[m,n] = h[m,n] > 0 ? sum_k dy[m,k] W[k,n] : 0
 * with W the nn.Linear weight [K = out, N = in] as it lies in memory and h the forward activations (bf16, same shape and
 * leading dimension as dx): the relu backward folded into the epilogue of the input-gradient product (bf16 throughput mode). */
int fira_gemm_bf16_tc_dx_relu(const void* dy, long lddy, const void* W, long ldw, void* dx, long lddx, const void* h,
                              int M, int N, int K, void* stream);

/* Debugging aid (tools/gemm_probe.py): with a device buffer of >= 16 uint64 set, CTA (0,0,0) of every following
 * fira_gemm_bf16_tc launch stamps %globaltimer at its phase boundaries (entry, prologue, dependency wait, TMA issued,
 * first stage landed, MMAs issued, accumulator ready, stores issued, exit); NULL switches it off (the default). */
int fira_debug_set_probe(void* probe);

/* ---- embeddings -------------------------------------------------------------------------------
 * Encoder node features in segment-major order (all code rows, all sub-token rows, all AST/edit
 * rows): emb[sou]+PE | emb[sub_token] | ast_emb[ast_change]     (gnn_transformer.py:46-52,58).
 * out_code holds rows [0, B*n_code); out_rest is indexed by the GLOBAL row (rows >= B*n_code). */
int fira_embed_nodes_fwd(const int* sou, const int* sub_token, const int* ast_change, const float* emb,
                         const float* ast_emb, const float* pos_table, void* out_code, void* out_rest, int B,
                         int n_code, int n_sub, int n_ast, int dim, int dtype, void* stream);
/* The same with an explicit position per code row (packed batches, fira_icse_b200/packed.py: `pos` = index of the token
 * inside its commit; B = 1, n_code / n_sub / n_ast = the padded row counts of the three segments). */
int fira_embed_nodes_pos_fwd(const int* sou, const int* pos, const int* sub_token, const int* ast_change,
                             const float* emb, const float* ast_emb, const float* pos_table, void* out_code,
                             void* out_rest, int B, int n_code, int n_sub, int n_ast, int dim, int dtype, void* stream);
/* Zero the segment-padding rows of a packed batch's [Rc + Rs, ld] memory-row matrix (off = its [3][B+1] row offsets). */
int fira_zero_pad_rows(void* x, long ld, int width, const int* off, int B, int Rc, int Rs, int dtype, void* stream);
int fira_embed_nodes_bwd(const int* sou, const int* sub_token, const int* ast_change, const void* d_code,
                         const void* d_rest, float* d_emb, float* d_ast_emb, int B, int n_code, int n_sub, int n_ast,
                         int dim, int dtype, void* stream);
/* Decoder input: dec_emb[tar] + PE[t]  (gnn_transformer.py:110-113). */
int fira_embed_rows_fwd(const int* ids, const float* emb, const float* pos_table, void* out, long rows, int period,
                        int dim, int dtype, void* stream);
int fira_embed_rows_bwd(const int* ids, const void* d_out, float* d_emb, long rows, int dim, int dtype, void* stream);

/* ---- LN(dropout(z) + resid)  (gnn_transformer.py:83,161,174,205) -------------------------------
 * rows < split are written to outA[row], the others to outB[row] (lets a GCN layer hand its code
 * rows to the next Combination without a torch.cat / slice copy). */
int fira_ln_residual_fwd(const void* z, const void* resid, const float* gamma, const float* beta, void* outA,
                         void* outB, long split, float* mean, float* rstd, long rows, int dim, float p_drop,
                         uint64_t seed, const uint64_t* seed_ctr, uint32_t stream_id, int dtype, void* stream);
int fira_ln_residual_bwd(const void* d_outA, const void* d_outB, long split, const void* z, const void* resid,
                         const float* mean, const float* rstd, const float* gamma, void* d_z, void* d_resid,
                         int d_resid_accum, float* d_gamma, float* d_beta, long rows, int dim, float p_drop,
                         uint64_t seed, const uint64_t* seed_ctr, uint32_t stream_id, int dtype, void* stream);

/* ---- Combination gate (combination_layer.py:7-17): c = v + sigmoid(q*(k-v)/sqrt(d_head))*(k-v),
 *      dropout; qk = [q | k] per row, v = vtab[mark[row]] (4 x dim table = Linear(mark_embedding)). */
int fira_comb_gate_fwd(const void* qk, long ld_qk, const float* vtab, const int* mark, void* out, long rows, int dim,
                       int d_head, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t stream_id, int dtype, void* stream);
int fira_comb_gate_bwd(const void* qk, long ld_qk, const float* vtab, const int* mark, const void* d_out, void* d_qk,
                       float* d_vtab, long rows, int dim, int d_head, float p_drop, uint64_t seed, const uint64_t* seed_ctr,
                       uint32_t stream_id, int dtype, void* stream);

/* The same gate with an arbitrary per-row `value` tensor (the stand-alone Combination.forward(query, key, value) /
 * CombinationLayer.forward of gnn_transformer.py:192-205, combination_layer.py:7-17): q, k, v, out are [rows, dim]. */
int fira_comb_gate3_fwd(const void* q, const void* k, const void* v, void* out, long rows, int dim, int d_head,
                        float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t stream_id, int dtype, void* stream);
int fira_comb_gate3_bwd(const void* q, const void* k, const void* v, const void* d_out, void* d_q, void* d_k, void* d_v,
                        long rows, int dim, int d_head, float p_drop, uint64_t seed, const uint64_t* seed_ctr,
                        uint32_t stream_id, int dtype, void* stream);

/* d[i] = h[i] > 0 ? d[i] : 0  (backward of the FeedForward relu, gnn_transformer.py:172). */
int fira_relu_bwd(const void* h, void* d, long n, int dtype, void* stream);
/* out[n] += sum_m w[m] * x[m,n]  (w == NULL -> 1): bias gradients. */
int fira_colsum(const void* x, long ld, long M, int N, const float* row_weight, float* out, int dtype, void* stream);

/* ---- optimizer step (run_model.py:101-109: `optimizer.step()` of torch.optim.Adam(lr), no weight decay / amsgrad) over
 *      ONE flat fp32 parameter buffer (fira_icse_b200/optim.py re-homes the model's parameters in it), one launch:
 *        g' = g / *grad_scale (grad_scale NULL -> 1);  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
 *        p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),  t = *step (device fp32 scalar, the caller has
 *        already incremented it);  p_bf16 (may be NULL) receives the bf16 copy of the updated parameters (the GEMM
 *        operands of the throughput mode).  n: multiple of 8; all buffers 16-byte aligned. */
int fira_adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, long n, float lr, float beta1, float beta2,
                   float eps, const float* step, const float* grad_scale, void* stream);
/* y = bf16(x), n a multiple of 8 (refresh of the bf16 parameter mirror after parameters were set from outside). */
int fira_cast_bf16(const float* x, void* y, long n, void* stream);

/* memory = cat(code rows, sub-token rows) per commit (Model.py:48) and its adjoint. */
int fira_pack_memory(const void* code, const void* rest, void* mem, int B, int n_code, int n_sub, int dim, int dtype,
                     void* stream);
int fira_unpack_memory(const void* d_mem, void* d_code, void* d_rest, int B, int n_code, int n_sub, int n_ast, int dim,
                       int dtype, void* stream);

/* ---- graph: dense [B,N,N] adjacency (Dataset.py:340 toarray(), any strides) -> packed CSR.
 * Two calls because the caller owns the buffers: count (+ exclusive scan into rowptr[B*N+1]),
 * read rowptr[B*N] to size col/val, then fill. */
int fira_csr_count_dense(const void* edge, int edge_dtype, long stride_b, long stride_i, long stride_j, int B, int N,
                         int* counts, int* rowptr, void* stream);
int fira_csr_fill_dense(const void* edge, int edge_dtype, long stride_b, long stride_i, long stride_j, int B, int N,
                        const int* rowptr, int* col, float* val, void* stream);
int fira_csr_rowsum(const int* rowptr, const float* val, int B, int n_code, int n_sub, int n_ast, float* out,
                    void* stream);
/* The GNN scatter: y = A x (+ addend), replacing torch.bmm(edge.float(), x) (gnn_transformer.py:80). */
int fira_gcn_aggregate(const int* rowptr, const int* col, const float* val, const void* x, const void* addend,
                       void* y, int B, int n_code, int n_sub, int n_ast, int dim, int dtype, void* stream);

/* ---- fused GCN layer, bf16 throughput mode (gnn_transformer.py:74-86 as ONE kernel per direction): gather the
 *      neighbour rows into the shared-memory A tile -> tcgen05.mma with the merged weight -> epilogue out of TMEM.
 *      The CSR is in BUFFER order: rowptr_rows[r] indexes the rows of the node buffer, col_rows are buffer rows
 *      (fira_csr_to_rows converts the (graph, node)-ordered CSR; counts is an int32[rows] workspace).
 *   fwd:  z = (A h) w_merged^T + rowsum(A) (x) c1 + bias ;  out = LN(dropout(z) + h)   (w_merged = fc2.W fc1.W [out,in]
 *         bf16, c1 = fc2.W fc1.b, bias = fc2.b; rows < split -> outA[row], the others -> outB[row]; mean/rstd for
 *         fira_ln_residual_bwd; the dropout mask is the one fira_ln_residual_fwd/bwd draw for (seed, stream_id)).
 *   bwd:  agg_dz = A^T dz (kept: d(w_merged) = agg_dz^T h, d(c1) = colsum(agg_dz)) ;  d_h = agg_dz w_merged + d_resid
 *         (w_merged_t = w_merged^T as a row-major [in,out] bf16 matrix; rowptr/col/val of A^T, = A when symmetric). */
int fira_csr_to_rows(const int* rowptr, const int* col, const float* val, int B, int n_code, int n_sub, int n_ast,
                     int* counts, int* rowptr_rows, int* col_rows, float* val_rows, void* stream);
int fira_gcn_layer_fwd(const int* rowptr_rows, const int* col_rows, const float* val_rows, const void* h,
                       const void* w_merged, const float* bias, const float* c1, const float* gamma, const float* beta,
                       void* z, void* outA, void* outB, long split, float* mean, float* rstd, long rows, int dim,
                       float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t stream_id, void* stream);
int fira_gcn_layer_bwd(const int* rowptr_rows_t, const int* col_rows_t, const float* val_rows_t, const void* d_z,
                       const void* w_merged_t, const void* d_resid, void* agg_dz, void* d_h, long rows, int dim,
                       void* stream);

/* ---- attention core (gnn_transformer.py:144-156); stats = (row max, row sum) [B,H,Lq,2];
 *      backward also takes the forward output `ctx` (same layout as d_ctx): delta = dO . O. */
int fira_attn_fwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                  const unsigned char* key_mask, int causal, void* ctx, long ldo, float* stats, int B, int H, int Lq,
                  int Lk, int d_head, int dtype, void* stream);
int fira_attn_bwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                  const unsigned char* key_mask, int causal, const void* ctx, const void* d_ctx, long ldo,
                  const float* stats, void* dq, long lddq, void* dk, long lddk, void* dv, long lddv, int B, int H,
                  int Lq, int Lk, int d_head, int dtype, void* stream);

/* Cross-attention of a PACKED batch (fira_icse_b200/packed.py): the keys / values of commit b are two row ranges of
 * k / v, ranges[b] = {first row, rows, first row, rows} (code rows, sub-token rows; GLOBAL row ids, kv_rows = rows of
 * k / v), key_mask [B, mask_pitch] over the commit's own key positions (NULL: all valid), mask_pitch >= rows of any
 * commit; max_chunks = an upper bound the caller guarantees on ceil(rows0 / 128) + ceil(rows1 / 128) of any commit
 * (fira_host_packed_dims reports it; <= 3 lets bf16 run on the tcgen05 kernels).  Rows of dk / dv outside every range
 * are not written (fira_zero_pad_rows clears the segment padding). */
int fira_attn_packed_fwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv, const int* ranges,
                         long kv_rows, const unsigned char* key_mask, int mask_pitch, int max_chunks, void* ctx, long ldo,
                         float* stats, int B, int H, int Lq, int d_head, int dtype, void* stream);
int fira_attn_packed_bwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv, const int* ranges,
                         long kv_rows, const unsigned char* key_mask, int mask_pitch, int max_chunks, const void* ctx,
                         const void* d_ctx, long ldo, const float* stats, void* dq, long lddq, void* dk, long lddk,
                         void* dv, long lddv, int B, int H, int Lq, int d_head, int dtype, void* stream);

/* ---- CopyNet scores (Model.py:17-18): sc[b,t,s] = b_res + w_res . tanh(src[b,s] + tgt[b,t]).
 *      src_mask [B,S] / row_mask [B*T] (optional, 1 = compute): positions the caller will mask anyway. */
int fira_copy_scores_fwd(const void* src_proj, const void* tgt_proj, const float* w_res, const float* b_res,
                         const unsigned char* src_mask, const unsigned char* row_mask, float* scores,
                         int B, int T_len, int S, int dim, int dtype, void* stream);
int fira_copy_scores_bwd(const void* src_proj, const void* tgt_proj, const float* w_res, const float* d_scores,
                         const unsigned char* row_active, void* d_src_proj, float* d_tgt_proj, float* d_w_res,
                         float* d_b_res, int B, int T_len, int S, int dim, int dtype, void* stream);

/* The same for a PACKED batch: the source rows of commit b are the two row ranges ranges[b] of src_proj (global rows);
 * scores stay [B, T, S] over the commit's own memory positions (S = mask pitch); d_src_proj rows outside the ranges
 * are not written. */
int fira_copy_scores_packed_fwd(const void* src_proj, const void* tgt_proj, const float* w_res, const float* b_res,
                                const int* ranges, const unsigned char* src_mask, const unsigned char* row_mask,
                                float* scores, int B, int T_len, int S, int dim, int dtype, void* stream);
int fira_copy_scores_packed_bwd(const void* src_proj, const void* tgt_proj, const float* w_res, const float* d_scores,
                                const unsigned char* row_active, const int* ranges, void* d_src_proj, float* d_tgt_proj,
                                float* d_w_res, float* d_b_res, int B, int T_len, int S, int dim, int dtype, void* stream);

/* ---- dual-copy mixture, loss and argmax (Model.py:54-86).  stats: 8 floats per row
 *      (vmax, vsum, cmax, csum, g0, g1, p_label, 0).  argmax_out may be NULL (training).
 *      logits / d_logits: 16-byte aligned, ld_logits a multiple of 8 (rows are read / written 8 elements at a time). */
int fira_pointer_mix_nll_fwd(const void* logits, long ld_logits, const float* copy_scores, const float* gate_logits,
                             const unsigned char* mem_mask, const int* label, float* stats, float* nll,
                             int* argmax_out, long rows, int T_len, int V, int S, int dtype, void* stream);
int fira_pointer_mix_nll_bwd(const void* logits, long ld_logits, const float* copy_scores,
                             const unsigned char* mem_mask, const int* label, const float* stats,
                             const float* upstream, void* d_logits, float* d_copy_scores, float* d_gate_logits,
                             unsigned char* row_active, long rows, int T_len, int V, int S, int dtype, void* stream);

/* ---- HOST-side batch preparation (CPU only: every pointer below is HOST memory, there is no stream).
 *
 * fira_host_build_adjacency: the commit graph of Dataset.py:220-294 + process_edge (Dataset.py:346-357).
 *   Relations are int32 pair lists [n,2] exactly as stored in DataSet/edge_*.json: (edit c, code j),
 *   (edit c, AST a), (AST a, code j), (AST a, AST b); code_sub = (code j, sub-token k) pairs
 *   (Dataset.py:173-192, 255-259); n_diff = diff tokens before padding (the sequential chain covers
 *   <start> t1..tn <eos>, Dataset.py:263-266); n_ast = AST nodes (edit nodes follow them).
 *   Node ids: code j -> j+1, sub-token k -> diff_len+k, AST a -> diff_len+sub_len+a, edit c -> ...+n_ast+c;
 *   pairs whose code id reaches diff_len are dropped (Dataset.py:228,243).  Output: undirected,
 *   de-duplicated, self loop on every node, in CSR order: deg[n_nodes], col[nnz], val[nnz] =
 *   1/sqrt(deg_row)/sqrt(deg_col) in float64 (Dataset.py:277-291).  *nnz_out is set even when cap is
 *   too small (FIRA_ERR_SHAPE).
 *
 * fira_host_batch_dims: segment lengths a batch needs after dropping the padding ALL its commits share:
 *   dims[3] = {c0, c1, c2} = position after the last non-zero id over commits index[0..batch) of the
 *   code / sub-token / AST+edit id tables, rounded up to mult_* (mult <= 0: keep the full length).
 *
 * fira_host_gather_batch: the loader step (Dataset.py:336-343 __getitem__ + default collate, minus the dense
 *   float64 toarray()): gathers commits index[0..batch) from the packed split arrays (int32 id tables
 *   [n, len], deg uint8 [n, n_nodes], col int16 / val float64 concatenated, edge_ptr int64 [n+1]) into
 *   caller-owned staging buffers (pinned memory): int64 id tensors [batch, c*] written with row length
 *   dims[0..2] (from fira_host_batch_dims, or any larger lengths up to the full 210/160/280), batch CSR
 *   rowptr int32 [batch*(c0+c1+c2)+1], col int32, val fp32.  The nodes cut away are isolated self loops
 *   (Dataset.py:271-275); sub-token copy labels (Dataset.py:213) shift by the removed code padding.
 *   Fails if a commit has a real id or a neighbour beyond dims. */
int fira_host_build_adjacency(const int* change_code, int n_change_code, const int* change_ast, int n_change_ast,
                              const int* ast_code, int n_ast_code, const int* ast_ast, int n_ast_ast,
                              const int* code_sub, int n_code_sub, int n_diff, int n_ast, int diff_len, int sub_len,
                              int ast_change_len, int* deg, int* col, double* val, int cap, int* nnz_out);
int fira_host_batch_dims(const int* sou, const int* sub_token, const int* ast_change, const long* index, int batch,
                         int diff_len, int sub_len, int ast_change_len, int mult_code, int mult_sub, int mult_ast,
                         int* dims);
int fira_host_gather_batch(const int* sou, const int* tar, const int* mark, const int* ast_change,
                           const int* tar_label, const int* sub_token, const unsigned char* deg, const short* col,
                           const double* val, const long* edge_ptr, const long* index, int batch, int diff_len,
                           int sub_len, int ast_change_len, int msg_len, int vocab_size, const int* dims,
                           long* o_sou, long* o_tar, long* o_mark, long* o_ast_change, long* o_tar_label,
                           long* o_sub_token, int* o_rowptr, int* o_col, float* o_val, long edge_cap, int* nnz_out);

/* fira_host_packed_dims / fira_host_gather_packed: the PER-COMMIT packed batch (replaces the
 *   fixed 210/160/280 padding of Dataset.py:80-94).  Per commit and segment only the positions up to the last non-zero
 *   id are kept; node rows are segment-major and ragged: [code rows of all commits | pad][sub-token rows | pad]
 *   [AST/edit rows | pad].  packed_dims -> dims[6] = {code rows, sub rows, AST rows, max memory rows of a commit, nnz,
 *   max over commits of ceil(code rows / 128) + ceil(sub rows / 128) = the key chunks fira_attn_packed_* needs}.
 *   gather_packed: pad_dims[4] = {Rc, Rs, Ra, S} buffer sizes (>= dims, bucketed by the caller); writes int32 node ids
 *   (o_code, o_mark, o_pos = position in the commit for the positional encoding; o_sub; o_ast), o_off[3][batch+1]
 *   (row offsets of each commit inside its segment), o_ranges[batch][4] = {first code row, code rows, first sub row
 *   (global), sub rows}, o_mem_mask[batch][S], the decoder input o_tar[batch][msg_len] + o_tar_mask, the SHIFTED labels
 *   (Model.py:71-79) with copy labels renumbered to the commit's own memory rows (V + m), and the adjacency as a CSR in
 *   buffer order with global column ids (what fira_gcn_layer_fwd / fira_gcn_aggregate(B=1) consume). */
int fira_host_packed_dims(const int* sou, const int* sub_token, const int* ast_change, const unsigned char* deg,
                          const long* index, int batch, int diff_len, int sub_len, int ast_change_len, int* dims);
int fira_host_gather_packed(const int* sou, const int* tar, const int* mark, const int* ast_change,
                            const int* tar_label, const int* sub_token, const unsigned char* deg, const short* col,
                            const double* val, const long* edge_ptr, const long* index, int batch, int diff_len,
                            int sub_len, int ast_change_len, int msg_len, int vocab_size, const int* pad_dims,
                            int* o_code, int* o_mark, int* o_pos, int* o_sub, int* o_ast, int* o_off, int* o_ranges,
                            unsigned char* o_mem_mask, int* o_tar, int* o_label, unsigned char* o_tar_mask,
                            int* o_rowptr, int* o_col, float* o_val, long edge_cap, int* nnz_out);

#ifdef __cplusplus
}
#endif
#endif /* FIRA_B200_H_ */
