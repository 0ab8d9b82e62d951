// Output head: CopyNet pointer scores (Model.py:15-20) and the dual-copy mixture / loss / argmax
// (Model.py:54-86).  Both are bandwidth/SFU kernels (no contraction worth a tensor core):
//   copy scores : sc[b,t,s] = b_res + sum_d w_res[d] * tanh(src[b,s,d] + tgt[b,t,d])
//                 the reference materialises the [B,30,370,256] tanh tensor (1.9 GB at B=170); here the
//                 30 target rows of a commit sit in shared memory and each warp streams source rows.
//   mixture     : p = [g0 * softmax(vocab logits) || g1 * softmax(masked copy scores)],
//                 logp = log(clamp(p, 1e-10, 1)), nll at the shifted label, argmax for 'dev'/'test'.
//                 Row-wise max/sum are warp/CTA reductions; the 25,020-wide distribution is never stored.
#include "common.cuh"
#include "fira_b200.h"

namespace {

constexpr int D = 256;
constexpr int TMAX = 32;   // tar_len 30 (run_model.py:32) padded

// source row of (commit b, memory position s): padded batches b*S + s; packed batches (ranges[b] = {first code row,
// code rows, first sub-token row, sub-token rows}) the s-th row of the two ranges, -1 beyond them
__device__ __forceinline__ long src_row(const int* __restrict__ ranges, int b, int S, int s) {
  if (!ranges) return (long)b * S + s;
  const int* r = ranges + 4 * b;
  if (s < r[1]) return r[0] + s;
  s -= r[1];
  return s < r[3] ? (long)(r[2] + s) : -1;
}

// ------------------------------------------------------------------ copy scores forward
template <typename T>
__global__ void __launch_bounds__(256) copy_scores_fwd_kernel(const T* __restrict__ src, const T* __restrict__ tgt,
                                                              const float* __restrict__ w_res,
                                                              const float* __restrict__ b_res_p,
                                                              const unsigned char* __restrict__ src_mask,
                                                              const unsigned char* __restrict__ row_mask,
                                                              const int* __restrict__ ranges,
                                                              float* __restrict__ sc, int B, int Tn, int S) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  // src_mask / row_mask (optional): padded source positions are overwritten with -1e9 by the mixture kernel
  // (Model.py:61) and target rows without a label never reach the loss -- both are skipped (score 0 written).
  const float b_res = *b_res_p;
  __shared__ __align__(16) float tg[TMAX][D];
  __shared__ __align__(16) float wr[D];
  __shared__ int active_t[TMAX];
  __shared__ int n_active;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows_per_cta = 32;
  const int s0 = blockIdx.x * rows_per_cta;
  const int s_end = min(S, s0 + rows_per_cta);
  // target rows that need scores (training: only rows whose label is a copy label).  The others get their zeros by
  // ROW segments (one 128-byte store per warp and row) instead of one scattered 4-byte store per (source row, target row)
  // out of the compute loop -- ~27 of 30 target rows of a commit are inactive.
  if (threadIdx.x == 0) {
    int n = 0;
    for (int t = 0; t < Tn; ++t) if (!row_mask || row_mask[(long)b * Tn + t]) active_t[n++] = t;
    n_active = n;
  }
  for (int idx = threadIdx.x; idx < D; idx += blockDim.x) wr[idx] = w_res[idx];
  __syncthreads();
  const int na = n_active;
  for (int idx = threadIdx.x; idx < na * D; idx += blockDim.x) {
    const int t = active_t[idx / D];
    tg[t][idx % D] = Act<T>::ld(tgt + ((long)b * Tn + t) * D + idx % D);
  }
  if (row_mask)
    for (int t = warp; t < Tn; t += 8)
      if (row_mask[(long)b * Tn + t] == 0 && s0 + lane < s_end) sc[((long)b * Tn + t) * S + s0 + lane] = 0.f;
  __syncthreads();
  float w[8];
  Act<float>::load8(wr + lane * 8, w);
  for (int s = s0 + warp; s < s_end; s += 8) {
    const long srow = src_row(ranges, b, S, s);
    if (srow < 0 || (src_mask && src_mask[(long)b * S + s] == 0)) {
      for (int a = lane; a < na; a += 32) sc[((long)b * Tn + active_t[a]) * S + s] = 0.f;
      continue;
    }
    float x[8];
    Act<T>::load8(src + srow * D + lane * 8, x);
    for (int a = 0; a < na; ++a) {
      const int t = active_t[a];
      float y[8];
      Act<float>::load8(&tg[t][lane * 8], y);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(w[i], tanhf(x[i] + y[i]), acc);
      acc = warp_sum(acc);
      if (lane == 0) sc[((long)b * Tn + t) * S + s] = acc + b_res;
    }
  }
}

// backward: only (b,t) rows flagged active carry gradient (rows whose label is a copy label).
//   d_src[b,s,:] = sum_t g[b,t,s] w (1 - th^2),  d_tgt[b,t,:] = sum_s (same),  d_w = sum g th,  d_b = sum g
template <typename T>
__global__ void __launch_bounds__(256) copy_scores_bwd_kernel(const T* __restrict__ src, const T* __restrict__ tgt,
                                                              const float* __restrict__ w_res,
                                                              const float* __restrict__ d_sc,
                                                              const unsigned char* __restrict__ row_active,
                                                              const int* __restrict__ ranges,
                                                              T* __restrict__ d_src, float* __restrict__ d_tgt,
                                                              float* __restrict__ d_w, float* __restrict__ d_b, int B,
                                                              int Tn, int S) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  extern __shared__ __align__(16) float dyn_smem[];
  float (*tg)[D] = reinterpret_cast<float (*)[D]>(dyn_smem);                    // [TMAX][D]
  float (*dtg)[D] = reinterpret_cast<float (*)[D]>(dyn_smem + TMAX * D);        // [TMAX][D]
  float* wr = dyn_smem + 2 * TMAX * D;                                          // [D]
  float (*red_w)[D] = reinterpret_cast<float (*)[D]>(dyn_smem + 2 * TMAX * D + D);  // [8][D]
  __shared__ int active_t[TMAX];
  __shared__ int n_active;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int t = 0; t < Tn; ++t) if (row_active[(long)b * Tn + t]) active_t[n++] = t;
    n_active = n;
  }
  for (int idx = threadIdx.x; idx < Tn * D; idx += blockDim.x) {
    tg[idx / D][idx % D] = Act<T>::ld(tgt + (long)b * Tn * D + idx);
    dtg[idx / D][idx % D] = 0.f;
  }
  for (int idx = threadIdx.x; idx < D; idx += blockDim.x) wr[idx] = w_res[idx];
  __syncthreads();
  const int na = n_active;
  float w[8], dw[8];
  Act<float>::load8(wr + lane * 8, w);
#pragma unroll
  for (int i = 0; i < 8; ++i) dw[i] = 0.f;
  float dbias = 0.f;
  const int rows_per_cta = 32;
  const int s_end = min(S, (int)(blockIdx.x + 1) * rows_per_cta);
  for (int s = blockIdx.x * rows_per_cta + warp; s < s_end; s += 8) {
    const long srow = src_row(ranges, b, S, s);
    if (srow < 0) continue;                            // beyond the commit's memory rows (packed batches)
    float x[8], dx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dx[i] = 0.f;
    if (na > 0) Act<T>::load8(src + srow * D + lane * 8, x);
    for (int a = 0; a < na; ++a) {
      const int t = active_t[a];
      const float g = d_sc[((long)b * Tn + t) * S + s];
      if (g == 0.f) continue;      // warp-uniform
      float y[8];
      Act<float>::load8(&tg[t][lane * 8], y);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float th = tanhf(x[i] + y[i]);
        const float term = g * w[i] * (1.f - th * th);
        dx[i] += term;
        dw[i] = fmaf(g, th, dw[i]);
        atomicAdd(&dtg[t][lane * 8 + i], term);
      }
      dbias += g;
    }
    Act<T>::store8(d_src + srow * D + lane * 8, dx);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red_w[warp][lane * 8 + i] = dw[i];
  __syncthreads();
  if (na > 0) {
    for (int idx = threadIdx.x; idx < D; idx += blockDim.x) {
      float sacc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) sacc += red_w[q][idx];
      atomicAdd(d_w + idx, sacc);
    }
    for (int a = 0; a < na; ++a) {
      const int t = active_t[a];
      for (int idx = threadIdx.x; idx < D; idx += blockDim.x) atomicAdd(d_tgt + ((long)b * Tn + t) * D + idx, dtg[t][idx]);
    }
    if (lane == 0 && dbias != 0.f) atomicAdd(d_b, dbias);
  }
}

// ------------------------------------------------------------------ mixture / loss / argmax
struct MaxSum { float m, s; };
__device__ __forceinline__ MaxSum ms_merge(MaxSum a, MaxSum b) {
  const float m = fmaxf(a.m, b.m);
  MaxSum r;
  r.m = m;
  r.s = (a.s == 0.f ? 0.f : a.s * expf(a.m - m)) + (b.s == 0.f ? 0.f : b.s * expf(b.m - m));
  return r;
}
__device__ __forceinline__ MaxSum ms_warp(MaxSum v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum u;
    u.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    u.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    v = ms_merge(v, u);
  }
  return v;
}
struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax am_better(ArgMax a, ArgMax b) {   // larger value, then smaller index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

// stats row layout (8 floats): vmax, vsum, cmax, csum, g0, g1, p_label, unused
template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ logits, long ldl,
                                                       const float* __restrict__ sc, const float* __restrict__ gate_logit,
                                                       const unsigned char* __restrict__ mem_mask,
                                                       const int* __restrict__ label, float* __restrict__ stats,
                                                       float* __restrict__ nll, int* __restrict__ argmax_out, int Tn,
                                                       int V, int S) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  __shared__ MaxSum sh_ms[8];
  __shared__ ArgMax sh_am[8];
  __shared__ float bc[8];
  const long row = blockIdx.x;
  const int b = (int)(row / Tn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* lrow = logits + row * ldl;
  const float* srow = sc + row * S;
  const unsigned char* mrow = mem_mask + (long)b * S;

  // pass 1: vocab max / sum-exp (online), copy max / sum-exp with the -1e9 fill (Model.py:61).
  // Training (no argmax wanted): a row's loss only needs the softmax its label lives in, so rows whose label is
  // padding or a copy label skip the 24,650-wide pass (and copy-label rows are the only ones that need `sc`).
  const int lab_row = label[row];
  const bool need_vocab = argmax_out != nullptr || (lab_row != 0 && lab_row < V);
  MaxSum v{-INFINITY, 0.f};
  if (need_vocab) {
    // 8 logits per 16-byte (bf16) / 32-byte (fp32) load; the running (max, sum) is rescaled once per group instead of
    // once per element (one 2-byte load and two expf per element made this kernel 61 us for 95 MB)
    const int V8 = V >> 3;
    for (int g = threadIdx.x; g < V8; g += blockDim.x) {
      float x[8];
      Act<T>::load8(lrow + (long)g * 8, x);
      float m8 = x[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) m8 = fmaxf(m8, x[i]);
      if (m8 > v.m) { v.s *= expf(v.m - m8); v.m = m8; }
#pragma unroll
      for (int i = 0; i < 8; ++i) v.s += expf(x[i] - v.m);
    }
    for (int j = V8 * 8 + threadIdx.x; j < V; j += blockDim.x) { MaxSum u{Act<T>::ld(lrow + j), 1.f}; v = ms_merge(v, u); }
  } else if (threadIdx.x == 0) v = MaxSum{0.f, 1.f};
  v = ms_warp(v);
  if (lane == 0) sh_ms[warp] = v;
  __syncthreads();
  if (warp == 0) { MaxSum u = lane < 8 ? sh_ms[lane] : MaxSum{-INFINITY, 0.f}; u = ms_warp(u); if (lane == 0) { bc[0] = u.m; bc[1] = u.s; } }
  __syncthreads();
  MaxSum c{-INFINITY, 0.f};
  for (int j = threadIdx.x; j < S; j += blockDim.x) { MaxSum u{mrow[j] ? srow[j] : kMaskFill, 1.f}; c = ms_merge(c, u); }
  c = ms_warp(c);
  if (lane == 0) sh_ms[warp] = c;
  __syncthreads();
  if (warp == 0) { MaxSum u = lane < 8 ? sh_ms[lane] : MaxSum{-INFINITY, 0.f}; u = ms_warp(u); if (lane == 0) { bc[2] = u.m; bc[3] = u.s; } }
  __syncthreads();
  const float vmax = bc[0], vsum = bc[1], cmax = bc[2], csum = bc[3];
  const float gl0 = gate_logit[row * 2], gl1 = gate_logit[row * 2 + 1];
  const float gm = fmaxf(gl0, gl1);
  const float e0 = expf(gl0 - gm), e1 = expf(gl1 - gm);
  const float g0 = e0 / (e0 + e1), g1 = e1 / (e0 + e1);

  if (threadIdx.x == 0) {
    const int lab = label[row];
    float p = 1.f;
    if (lab != 0) {
      if (lab < V) p = g0 * (expf(Act<T>::ld(lrow + lab) - vmax) / vsum);
      else {
        // a copy label beyond the (possibly loader-trimmed) source is never read out of bounds: it gets p = 0 -> the
        // clamp floor, no gradient -- what the reference computes for a label on a padded (masked) source position
        // (Model.py:61,69); beyond V+370 the reference's nll_loss raises instead (Model.py:81)
        const int s = lab - V;
        p = s < S ? g1 * (expf((mrow[s] ? srow[s] : kMaskFill) - cmax) / csum) : 0.f;
      }
    }
    float* st = stats + row * 8;
    st[0] = vmax; st[1] = vsum; st[2] = cmax; st[3] = csum; st[4] = g0; st[5] = g1; st[6] = p; st[7] = 0.f;
    // loss = -log(clamp(p, 1e-10, 1)), zeroed where label == 0 (Model.py:69,81-82)
    nll[row] = lab != 0 ? -logf(fminf(fmaxf(p, 1e-10f), 1.f)) : 0.f;
  }
  if (argmax_out) {
    // argmax over log(clamp(p)) of the concatenation, first index wins ties (Model.py:86)
    ArgMax best{-INFINITY, 0x7fffffff};
    const float iv = 1.f / vsum, ic = 1.f / csum;
    for (int j = threadIdx.x; j < V + S; j += blockDim.x) {
      float p;
      if (j < V) p = g0 * (expf(Act<T>::ld(lrow + j) - vmax) * iv);
      else { const int s = j - V; p = g1 * (expf((mrow[s] ? srow[s] : kMaskFill) - cmax) * ic); }
      ArgMax cand{logf(fminf(fmaxf(p, 1e-10f), 1.f)), j};
      best = am_better(best, cand);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax u{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = am_better(best, u);
    }
    if (lane == 0) sh_am[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      ArgMax r = sh_am[0];
      for (int q = 1; q < 8; ++q) r = am_better(r, sh_am[q]);
      argmax_out[row] = r.i;
    }
  }
}

// d(loss_sum)/d(logits, copy scores, gate logits); `upstream` is d(loss_sum) (a device scalar).
// Exactly one of the two softmaxes receives gradient per row (the picked element decides), rows with
// label == 0 or p outside [1e-10, 1] (clamp) receive none.
template <typename T>
__global__ void __launch_bounds__(256) head_bwd_kernel(const T* __restrict__ logits, long ldl,
                                                       const float* __restrict__ sc,
                                                       const unsigned char* __restrict__ mem_mask,
                                                       const int* __restrict__ label, const float* __restrict__ stats,
                                                       const float* __restrict__ upstream, T* __restrict__ d_logits,
                                                       float* __restrict__ d_sc, float* __restrict__ d_gate_logit,
                                                       unsigned char* __restrict__ row_active, int Tn, int V, int S) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  const long row = blockIdx.x;
  const int b = (int)(row / Tn);
  const float* st = stats + row * 8;
  const float vmax = st[0], vsum = st[1], cmax = st[2], csum = st[3], g0 = st[4], g1 = st[5], p = st[6];
  const int lab = label[row];
  const float up = *upstream;
  const bool live = lab != 0 && p >= 1e-10f && p <= 1.f;
  const bool vocab = live && lab < V;
  const bool copy = live && lab >= V;
  T* drow = d_logits + row * ldl;
  const T* lrow = logits + row * ldl;
  const int V8 = V >> 3;                         // 8 logits per vector load / store, scalar tail
  if (vocab) {
    const float iv = 1.f / vsum;
    for (int g = threadIdx.x; g < V8; g += blockDim.x) {
      float x[8];
      Act<T>::load8(lrow + (long)g * 8, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = up * (expf(x[i] - vmax) * iv - (g * 8 + i == lab ? 1.f : 0.f));
      Act<T>::store8(drow + (long)g * 8, x);
    }
    for (int j = V8 * 8 + threadIdx.x; j < V; j += blockDim.x) {
      float pv = expf(Act<T>::ld(lrow + j) - vmax) * iv;
      Act<T>::st(drow + j, up * (pv - (j == lab ? 1.f : 0.f)));
    }
  } else {
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int g = threadIdx.x; g < V8; g += blockDim.x) Act<T>::store8(drow + (long)g * 8, z);
    for (int j = V8 * 8 + threadIdx.x; j < V; j += blockDim.x) Act<T>::st(drow + j, 0.f);
  }
  const float* srow = sc + row * S;
  const unsigned char* mrow = mem_mask + (long)b * S;
  float* dsrow = d_sc + row * S;
  if (copy) {
    const float ic = 1.f / csum;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
      float pc = expf((mrow[j] ? srow[j] : kMaskFill) - cmax) * ic;
      // masked positions were overwritten by masked_fill -> no gradient reaches the raw score
      dsrow[j] = mrow[j] ? up * (pc - (j == lab - V ? 1.f : 0.f)) : 0.f;
    }
  } else {
    for (int j = threadIdx.x; j < S; j += blockDim.x) dsrow[j] = 0.f;
  }
  if (threadIdx.x == 0) {
    float d0 = 0.f, d1 = 0.f;
    if (vocab) { d0 = up * (g0 - 1.f); d1 = up * g1; }
    if (copy) { d0 = up * g0; d1 = up * (g1 - 1.f); }
    d_gate_logit[row * 2] = d0; d_gate_logit[row * 2 + 1] = d1;
    row_active[row] = copy ? 1 : 0;
  }
}

}  // namespace

#define DISPATCH_T(dtype, ...)                                                            \
  if ((dtype) == FIRA_F32) { using T = float; __VA_ARGS__ }                               \
  else if ((dtype) == FIRA_BF16) { using T = __nv_bfloat16; __VA_ARGS__ }                 \
  else { fira_set_error(FIRA_ERR_DTYPE, "unknown dtype %d", (int)(dtype)); return FIRA_ERR_DTYPE; }

extern "C" {

static int copy_scores_fwd_impl(const void* src_proj, const void* tgt_proj, const float* w_res, const float* b_res,
                                const unsigned char* src_mask, const unsigned char* row_mask, const int* ranges,
                                float* scores, int B, int T_len, int S, int dim, int dtype, void* stream) {
  FIRA_CHECK_ARG(dim == D, FIRA_ERR_SHAPE, "copy_scores_fwd: dim %d != 256", dim);
  FIRA_CHECK_ARG(T_len > 0 && T_len <= TMAX, FIRA_ERR_SHAPE, "copy_scores_fwd: T_len %d > %d", T_len, TMAX);
  if (B == 0 || S == 0) return FIRA_OK;
  dim3 grid((S + 31) / 32, B);
  DISPATCH_T(dtype, launch_k(copy_scores_fwd_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
      (const T*)src_proj, (const T*)tgt_proj, w_res, b_res, src_mask, row_mask, ranges, scores, B, T_len, S);)
  FIRA_CHECK_LAUNCH("fira_copy_scores_fwd");
  return FIRA_OK;
}

static int copy_scores_bwd_impl(const void* src_proj, const void* tgt_proj, const float* w_res, const float* d_scores,
                                const unsigned char* row_active, const int* ranges, void* d_src_proj, float* d_tgt_proj,
                                float* d_w_res, float* d_b_res, int B, int T_len, int S, int dim, int dtype, void* stream) {
  FIRA_CHECK_ARG(dim == D, FIRA_ERR_SHAPE, "copy_scores_bwd: dim %d != 256", dim);
  FIRA_CHECK_ARG(T_len > 0 && T_len <= TMAX, FIRA_ERR_SHAPE, "copy_scores_bwd: T_len %d > %d", T_len, TMAX);
  if (B == 0 || S == 0) return FIRA_OK;
  dim3 grid((S + 31) / 32, B);
  const int smem = (int)sizeof(float) * (2 * TMAX * D + D + 8 * D);
  cudaError_t e = dtype == FIRA_F32
      ? cudaFuncSetAttribute(copy_scores_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
      : cudaFuncSetAttribute(copy_scores_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { fira_set_error(FIRA_ERR_CUDA, "copy_scores_bwd attr: %s", cudaGetErrorString(e)); return FIRA_ERR_CUDA; }
  DISPATCH_T(dtype, launch_k(copy_scores_bwd_kernel<T>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, 
      (const T*)src_proj, (const T*)tgt_proj, w_res, d_scores, row_active, ranges, (T*)d_src_proj, d_tgt_proj, d_w_res,
      d_b_res, B, T_len, S);)
  FIRA_CHECK_LAUNCH("fira_copy_scores_bwd");
  return FIRA_OK;
}

int fira_copy_scores_fwd(const void* src_proj, const void* tgt_proj, const float* w_res, const float* b_res,
                         const unsigned char* src_mask, const unsigned char* row_mask, float* scores,
                         int B, int T_len, int S, int dim, int dtype, void* stream) {
  return copy_scores_fwd_impl(src_proj, tgt_proj, w_res, b_res, src_mask, row_mask, nullptr, scores, B, T_len, S, dim,
                              dtype, stream);
}

int fira_copy_scores_bwd(const void* src_proj, const void* tgt_proj, const float* w_res, const float* d_scores,
                         const unsigned char* row_active, void* d_src_proj, float* d_tgt_proj, float* d_w_res,
                         float* d_b_res, int B, int T_len, int S, int dim, int dtype, void* stream) {
  return copy_scores_bwd_impl(src_proj, tgt_proj, w_res, d_scores, row_active, nullptr, d_src_proj, d_tgt_proj, d_w_res,
                              d_b_res, B, T_len, S, dim, dtype, stream);
}

int fira_copy_scores_packed_fwd(const void* src_proj, const void* tgt_proj, const float* w_res, const float* b_res,
                                const int* ranges, const unsigned char* src_mask, const unsigned char* row_mask,
                                float* scores, int B, int T_len, int S, int dim, int dtype, void* stream) {
  FIRA_CHECK_ARG(ranges, FIRA_ERR_ARG, "copy_scores_packed_fwd: null ranges");
  return copy_scores_fwd_impl(src_proj, tgt_proj, w_res, b_res, src_mask, row_mask, ranges, scores, B, T_len, S, dim,
                              dtype, stream);
}

int fira_copy_scores_packed_bwd(const void* src_proj, const void* tgt_proj, const float* w_res, const float* d_scores,
                                const unsigned char* row_active, const int* ranges, void* d_src_proj, float* d_tgt_proj,
                                float* d_w_res, float* d_b_res, int B, int T_len, int S, int dim, int dtype, void* stream) {
  FIRA_CHECK_ARG(ranges, FIRA_ERR_ARG, "copy_scores_packed_bwd: null ranges");
  return copy_scores_bwd_impl(src_proj, tgt_proj, w_res, d_scores, row_active, ranges, d_src_proj, d_tgt_proj, d_w_res,
                              d_b_res, B, T_len, S, dim, dtype, stream);
}

int fira_pointer_mix_nll_fwd(const void* logits, long ld_logits, const float* copy_scores, const float* gate_logits,
                             const unsigned char* mem_mask, const int* label, float* stats, float* nll,
                             int* argmax_out, long rows, int T_len, int V, int S, int dtype, void* stream) {
  FIRA_CHECK_ARG(rows >= 0 && T_len > 0 && V > 0 && S > 0, FIRA_ERR_SHAPE, "pointer_mix_nll_fwd: shape");
  FIRA_CHECK_ARG(fira_aligned16(logits) && ld_logits % 8 == 0, FIRA_ERR_ALIGN,
                 "pointer_mix_nll_fwd: logits must be 16-byte aligned with a leading dimension that is a multiple of 8");
  if (rows == 0) return FIRA_OK;
  DISPATCH_T(dtype, launch_k(head_fwd_kernel<T>, dim3((unsigned)rows), dim3(256), 0, (cudaStream_t)stream, 
      (const T*)logits, ld_logits, copy_scores, gate_logits, mem_mask, label, stats, nll, argmax_out, T_len, V, S);)
  FIRA_CHECK_LAUNCH("fira_pointer_mix_nll_fwd");
  return FIRA_OK;
}

int fira_pointer_mix_nll_bwd(const void* logits, long ld_logits, const float* copy_scores,
                             const unsigned char* mem_mask, const int* label, const float* stats,
                             const float* upstream, void* d_logits, float* d_copy_scores, float* d_gate_logits,
                             unsigned char* row_active, long rows, int T_len, int V, int S, int dtype, void* stream) {
  FIRA_CHECK_ARG(rows >= 0 && T_len > 0 && V > 0 && S > 0, FIRA_ERR_SHAPE, "pointer_mix_nll_bwd: shape");
  FIRA_CHECK_ARG(fira_aligned16(logits) && fira_aligned16(d_logits) && ld_logits % 8 == 0, FIRA_ERR_ALIGN,
                 "pointer_mix_nll_bwd: logits / d_logits must be 16-byte aligned with a leading dimension that is a multiple of 8");
  if (rows == 0) return FIRA_OK;
  DISPATCH_T(dtype, launch_k(head_bwd_kernel<T>, dim3((unsigned)rows), dim3(256), 0, (cudaStream_t)stream, 
      (const T*)logits, ld_logits, copy_scores, mem_mask, label, stats, upstream, (T*)d_logits, d_copy_scores,
      d_gate_logits, row_active, T_len, V, S);)
  FIRA_CHECK_LAUNCH("fira_pointer_mix_nll_bwd");
  return FIRA_OK;
}

}  // extern "C"
