// Multi-head attention core (gnn_transformer.py:144-156), one CTA per (commit, head):
//     S = Q K^T / sqrt(d_head);  S[mask == 0] = -1e9;  P = softmax(S);  ctx = P V
// mask = key padding (and causal for decoder self-attention, gnn_transformer.py:117); no dropout on P.
//
// The problem is tiny (Lq = 30, Lk <= 370, d_head = 32) and mostly padding: on the DataSet only ~127 of the
// 370 memory rows are real.  exp(-1e9 - max) is exactly 0 in fp32, so masked keys contribute nothing --
// the kernels COMPACT the valid keys first (ballot scan into a shared index list) and only load / multiply
// those (3x less work on real data, bit-for-bit the same sums).  A fully masked row (all keys padded)
// keeps the reference's behaviour: uniform P over all keys, no gradient to the scores.
// Each warp handles two query rows at a time so every K / V shared-memory read feeds two rows.
// Forward saves (row max, row sum); backward recomputes P from them, gets delta = rowsum(P * dP) as
// dO . O (so keys can be processed in chunks of 64 with a small shared-memory footprint, 5 CTAs/SM)
// and writes zero gradients for the masked keys.
#include <stdlib.h>
#include "common.cuh"
#include "fira_b200.h"

// tcgen05 path of the bf16 mode (attention_tc.cu)
int fira_attn_tc_fwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                     const unsigned char* key_mask, const int* ranges, long kv_rows, int causal, void* ctx, long ldo,
                     float* stats, int B, int H, int Lq, int Lk, void* stream);
int fira_attn_tc_bwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                     const unsigned char* key_mask, const int* ranges, long kv_rows, int causal, const void* ctx,
                     const void* d_ctx, long ldo, const float* stats, void* dq, long lddq, void* dk, long lddk, void* dv,
                     long lddv, int B, int H, int Lq, int Lk, void* stream);
bool fira_attn_tc_eligible(int B, int H, int Lq, int Lk, int d_head, long ldk, long ldv, int max_chunks);

namespace {

// bf16 activations go to the tensor-core kernels unless FIRA_ATTN_TC=0 (A/B runs against the FFMA kernels below)
bool use_tc(int dtype, int B, int H, int Lq, int Lk, int d_head, long ldk, long ldv, int max_chunks) {
  if (dtype != FIRA_BF16 || !fira_attn_tc_eligible(B, H, Lq, Lk, d_head, ldk, ldv, max_chunks)) return false;
  const char* e = getenv("FIRA_ATTN_TC");      // read per call: an A/B switch, not cached process state
  return e ? atoi(e) != 0 : true;
}

constexpr int DH = 32;           // head dim (256 / 8)
constexpr int KPAD = DH + 1;     // conflict-free column reads of K/V tiles
constexpr int NWARPS = 8;
constexpr int NTHR = NWARPS * 32;
constexpr int KC = 128;          // keys per forward chunk
constexpr int KCB = 64;          // keys per backward chunk (42 KB shared memory -> 5 CTAs per SM)
constexpr int LQ_MAX = 32;       // tar_len 30 (run_model.py:32)

struct AttnArgs {
  const void* q; long ldq;       // row (b*Lq + t), head h at column h*32
  const void* k; long ldk;       // row (b*Lk + s)
  const void* v; long ldv;
  const unsigned char* key_mask; // [B, Lk], 1 = attend (may be NULL with ranges: all keys valid)
  const int* ranges;             // NULL, or [B][4] = {first row, rows, first row, rows} of k / v (packed batches)
  int causal;
  int B, H, Lq, Lk;
  float scale;
};

// 16-byte loads of a 32-wide head slice; row index taken from `rows_idx` (compacted keys) or identity
template <typename T>
__device__ __forceinline__ void load_rows(float* dst, const T* src, long ld, const int* rows_idx, int row0, int nrows) {
  constexpr int EPV = 16 / sizeof(T);
  constexpr int VPR = DH / EPV;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < nrows * VPR; idx += NTHR) {
    const int j = idx / VPR, c = idx % VPR;
    const int r = rows_idx ? rows_idx[row0 + j] : row0 + j;
    const uint4 raw = *reinterpret_cast<const uint4*>(src + (long)r * ld + c * EPV);
    float* o = dst + j * KPAD + c * EPV;
    if constexpr (sizeof(T) == 4) {
      o[0] = __uint_as_float(raw.x); o[1] = __uint_as_float(raw.y);
      o[2] = __uint_as_float(raw.z); o[3] = __uint_as_float(raw.w);
    } else {
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(hh[q]); o[2 * q] = f.x; o[2 * q + 1] = f.y; }
    }
  }
}

// valid-key compaction by warp 0: kidx[0..nv) = original indices of keys with mask == 1 (ascending).
// nv == 0 (every key padded): identity list of all keys and *filled = 1 (scores are all -1e9 -> uniform P).
// causal (self-attention, Lk = 30): no compaction -- a row whose every permitted key is padding must stay
// uniform over ALL keys like the reference, so padding is handled by the score mask instead.
__device__ __forceinline__ void compact_keys(const unsigned char* km, int Lk, int causal, int* kidx, int* nv_out,
                                             int* filled, const int* rg = nullptr) {
  // rg != NULL (packed batches): key position m of the commit lives in global row
  //   m < rg[1] ? rg[0] + m : rg[2] + (m - rg[1]),   m < rg[1] + rg[3];   kidx then holds GLOBAL rows
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int L = rg ? rg[1] + rg[3] : Lk;
    int n = 0;
    for (int s0 = 0; s0 < L && !causal; s0 += 32) {
      const int s = s0 + lane;
      const bool ok = s < L && (km == nullptr || km[s] != 0);
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) kidx[n + __popc(bal & ((1u << lane) - 1u))] = rg ? (s < rg[1] ? rg[0] + s : rg[2] + (s - rg[1])) : s;
      n += __popc(bal);
    }
    if (n == 0) {
      for (int s = lane; s < L; s += 32) kidx[s] = rg ? (s < rg[1] ? rg[0] + s : rg[2] + (s - rg[1])) : s;
      n = L;
      if (lane == 0) *filled = causal ? 0 : 1;
    } else if (lane == 0) *filled = 0;
    if (lane == 0) *nv_out = n;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ forward
// Keys are consumed in chunks of KC with a running (max, sum, output) per query row (online softmax), so the
// shared-memory footprint is ~48 KB whatever Lk is: 4 CTAs per SM instead of 1 for Lk = 370.
constexpr int PAIRS_MAX = (LQ_MAX / 2 + NWARPS - 1) / NWARPS;      // query-row pairs per warp (2)

template <typename T>
__global__ void __launch_bounds__(NTHR) attn_fwd_kernel(AttnArgs a, T* __restrict__ ctx, long ldo,
                                                        float* __restrict__ stats /* [B,H,Lq,2] */) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  extern __shared__ float smem[];
  __shared__ int nv_s, filled_s;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* Ks = smem;                          // [KC][KPAD]
  float* Vs = Ks + KC * KPAD;                // [KC][KPAD]
  float* Qs = Vs + KC * KPAD;                // [Lq][KPAD]
  float* Ps = Qs + a.Lq * KPAD;              // [NWARPS][2][KC]
  int* kidx = reinterpret_cast<int*>(Ps + NWARPS * 2 * KC);   // [Lk]
  const unsigned char* km = a.key_mask ? a.key_mask + (long)b * a.Lk : nullptr;
  const int* rg = a.ranges ? a.ranges + 4 * b : nullptr;
  const long kvb = rg ? 0 : (long)b * a.Lk;                    // with ranges kidx holds global rows
  compact_keys(km, a.Lk, a.causal, kidx, &nv_s, &filled_s, rg);
  const int nv = nv_s;
  const bool filled = filled_s != 0;
  load_rows(Qs, (const T*)a.q + (long)b * a.Lq * a.ldq + h * DH, a.ldq, nullptr, 0, a.Lq);
  float* P0 = Ps + (warp * 2 + 0) * KC;
  float* P1 = Ps + (warp * 2 + 1) * KC;
  float m0[PAIRS_MAX], m1[PAIRS_MAX], l0[PAIRS_MAX], l1[PAIRS_MAX], o0[PAIRS_MAX], o1[PAIRS_MAX];
#pragma unroll
  for (int i = 0; i < PAIRS_MAX; ++i) { m0[i] = m1[i] = -INFINITY; l0[i] = l1[i] = 0.f; o0[i] = o1[i] = 0.f; }

  for (int c0 = 0; c0 < nv; c0 += KC) {
    const int nc = min(KC, nv - c0);
    __syncthreads();                                       // previous chunk consumed; Qs visible
    load_rows(Ks, (const T*)a.k + kvb * a.ldk + h * DH, a.ldk, kidx, c0, nc);
    load_rows(Vs, (const T*)a.v + kvb * a.ldv + h * DH, a.ldv, kidx, c0, nc);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PAIRS_MAX; ++i) {
      const int t0 = (warp + i * NWARPS) * 2;
      if (t0 < a.Lq) {
        const int t1 = min(t0 + 1, a.Lq - 1);              // odd Lq: row duplicated, second result dropped
        float q0[DH], q1[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) { q0[d] = Qs[t0 * KPAD + d]; q1[d] = Qs[t1 * KPAD + d]; }
        float cm0 = -INFINITY, cm1 = -INFINITY;
        for (int j = lane; j < nc; j += 32) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int d = 0; d < DH; ++d) { const float kd = Ks[j * KPAD + d]; s0 = fmaf(q0[d], kd, s0); s1 = fmaf(q1[d], kd, s1); }
          const int ko = kidx[c0 + j];
          const bool pad = filled || (a.causal && km[ko] == 0);
          s0 = (pad || (a.causal && ko > t0)) ? kMaskFill : s0 * a.scale;
          s1 = (pad || (a.causal && ko > t1)) ? kMaskFill : s1 * a.scale;
          P0[j] = s0; P1[j] = s1;
          cm0 = fmaxf(cm0, s0); cm1 = fmaxf(cm1, s1);
        }
        const float nm0 = fmaxf(m0[i], warp_max(cm0)), nm1 = fmaxf(m1[i], warp_max(cm1));
        const float r0 = expf(m0[i] - nm0), r1 = expf(m1[i] - nm1);     // exp(-inf) = 0 on the first chunk
        float sum0 = 0.f, sum1 = 0.f;
        for (int j = lane; j < nc; j += 32) {
          const float e0 = expf(P0[j] - nm0), e1 = expf(P1[j] - nm1);
          P0[j] = e0; P1[j] = e1; sum0 += e0; sum1 += e1;
        }
        l0[i] = l0[i] * r0 + warp_sum(sum0);
        l1[i] = l1[i] * r1 + warp_sum(sum1);
        m0[i] = nm0; m1[i] = nm1;
        __syncwarp();
        float a0 = o0[i] * r0, a1 = o1[i] * r1;
        for (int j = 0; j < nc; ++j) { const float vv = Vs[j * KPAD + lane]; a0 = fmaf(P0[j], vv, a0); a1 = fmaf(P1[j], vv, a1); }
        o0[i] = a0; o1[i] = a1;
        __syncwarp();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PAIRS_MAX; ++i) {
    const int t0 = (warp + i * NWARPS) * 2;
    if (t0 < a.Lq) {
      const int t1 = min(t0 + 1, a.Lq - 1);
      Act<T>::st(ctx + ((long)b * a.Lq + t0) * ldo + h * DH + lane, o0[i] / l0[i]);
      if (t1 != t0) Act<T>::st(ctx + ((long)b * a.Lq + t1) * ldo + h * DH + lane, o1[i] / l1[i]);
      if (lane == 0 && stats) {
        float* st = stats + (((long)b * a.H + h) * a.Lq + t0) * 2;
        st[0] = m0[i]; st[1] = l0[i];
        if (t1 != t0) { st[2] = m1[i]; st[3] = l1[i]; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// dV = P^T dO, dP = dO V^T, dS = P * (dP - delta), delta = dO . O, dQ = scale dS K, dK = scale dS^T Q.
template <typename T>
__global__ void __launch_bounds__(NTHR) attn_bwd_kernel(AttnArgs a, const T* __restrict__ out, const T* __restrict__ d_ctx,
                                                        long ldo, const float* __restrict__ stats, T* __restrict__ dq,
                                                        long lddq, T* __restrict__ dk, long lddk, T* __restrict__ dv,
                                                        long lddv) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  extern __shared__ float smem[];
  __shared__ int nv_s, filled_s;
  __shared__ float delta_s[LQ_MAX], mx_s[LQ_MAX], inv_s[LQ_MAX];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PP = KCB + 1;
  float* Ks = smem;                          // [KCB][KPAD]
  float* Vs = Ks + KCB * KPAD;                // [KCB][KPAD]
  float* Qs = Vs + KCB * KPAD;                // [Lq][KPAD]
  float* Os = Qs + a.Lq * KPAD;              // [Lq][KPAD]  dO
  float* Pm = Os + a.Lq * KPAD;              // [Lq][PP]    P
  float* Sm = Pm + a.Lq * PP;                // [Lq][PP]    dS
  int* kidx = reinterpret_cast<int*>(Sm + a.Lq * PP);   // [Lk]
  const unsigned char* km = a.key_mask ? a.key_mask + (long)b * a.Lk : nullptr;
  const int* rg = a.ranges ? a.ranges + 4 * b : nullptr;
  const long kvb = rg ? 0 : (long)b * a.Lk;
  compact_keys(km, a.Lk, a.causal, kidx, &nv_s, &filled_s, rg);
  const int nv = nv_s;
  const bool filled = filled_s != 0;
  // delta_t = dO_t . O_t  (== sum_s P dP), row statistics
  load_rows(Qs, out + (long)b * a.Lq * ldo + h * DH, ldo, nullptr, 0, a.Lq);       // O, temporarily in Qs
  load_rows(Os, d_ctx + (long)b * a.Lq * ldo + h * DH, ldo, nullptr, 0, a.Lq);
  __syncthreads();
  for (int t = warp; t < a.Lq; t += NWARPS) {
    const float x = warp_sum(Qs[t * KPAD + lane] * Os[t * KPAD + lane]);
    if (lane == 0) {
      delta_s[t] = x;
      const float* st = stats + (((long)b * a.H + h) * a.Lq + t) * 2;
      mx_s[t] = st[0]; inv_s[t] = 1.f / st[1];
    }
  }
  __syncthreads();
  load_rows(Qs, (const T*)a.q + (long)b * a.Lq * a.ldq + h * DH, a.ldq, nullptr, 0, a.Lq);
  float dqa[(LQ_MAX + NWARPS - 1) / NWARPS];   // dQ[t][lane] for this warp's rows t = warp + 8*i
#pragma unroll
  for (int i = 0; i < (LQ_MAX + NWARPS - 1) / NWARPS; ++i) dqa[i] = 0.f;

  for (int c0 = 0; c0 < nv; c0 += KCB) {
    const int nc = min(KCB, nv - c0);
    __syncthreads();                                     // previous chunk fully consumed (and Qs loaded)
    load_rows(Ks, (const T*)a.k + kvb * a.ldk + h * DH, a.ldk, kidx, c0, nc);
    load_rows(Vs, (const T*)a.v + kvb * a.ldv + h * DH, a.ldv, kidx, c0, nc);
    __syncthreads();
    // ---- phase A: P, dS for the warp's query rows; dQ accumulation
#pragma unroll
    for (int i = 0; i < (LQ_MAX + NWARPS - 1) / NWARPS; ++i) {
      const int t = warp + i * NWARPS;
      if (t < a.Lq) {
        float qr[DH], orr[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) { qr[d] = Qs[t * KPAD + d]; orr[d] = Os[t * KPAD + d]; }
        const float mx = mx_s[t], inv = inv_s[t], dl = delta_s[t];
        for (int j = lane; j < nc; j += 32) {
          float sc = 0.f, dp = 0.f;
#pragma unroll
          for (int d = 0; d < DH; ++d) { sc = fmaf(qr[d], Ks[j * KPAD + d], sc); dp = fmaf(orr[d], Vs[j * KPAD + d], dp); }
          const bool masked = filled || (a.causal && (kidx[c0 + j] > t || km[kidx[c0 + j]] == 0));
          sc = masked ? kMaskFill : sc * a.scale;
          const float pr = expf(sc - mx) * inv;
          Pm[t * PP + j] = pr;
          Sm[t * PP + j] = masked ? 0.f : pr * (dp - dl);      // masked_fill blocks the gradient
        }
        __syncwarp();
        float g = 0.f;
        for (int j = 0; j < nc; ++j) g = fmaf(Sm[t * PP + j], Ks[j * KPAD + lane], g);
        dqa[i] += g;
      }
    }
    __syncthreads();
    // ---- phase B: dK, dV rows of this chunk (written at the keys' original positions)
    for (int j = warp; j < nc; j += NWARPS) {
      float gk = 0.f, gv = 0.f;
      for (int t = 0; t < a.Lq; ++t) {
        gk = fmaf(Sm[t * PP + j], Qs[t * KPAD + lane], gk);
        gv = fmaf(Pm[t * PP + j], Os[t * KPAD + lane], gv);
      }
      const long row = kvb + kidx[c0 + j];
      Act<T>::st(dk + row * lddk + h * DH + lane, gk * a.scale);
      Act<T>::st(dv + row * lddv + h * DH + lane, gv);
    }
  }
#pragma unroll
  for (int i = 0; i < (LQ_MAX + NWARPS - 1) / NWARPS; ++i) {
    const int t = warp + i * NWARPS;
    if (t < a.Lq) Act<T>::st(dq + ((long)b * a.Lq + t) * lddq + h * DH + lane, dqa[i] * a.scale);
  }
  // masked keys receive exactly zero gradient
  if (!filled && !a.causal && km) {
    const int L = rg ? rg[1] + rg[3] : a.Lk;
    for (int s = warp; s < L; s += NWARPS) {
      if (km[s] == 0) {
        const long row = rg ? (s < rg[1] ? rg[0] + s : rg[2] + (s - rg[1])) : (long)b * a.Lk + s;
        Act<T>::st(dk + row * lddk + h * DH + lane, 0.f);
        Act<T>::st(dv + row * lddv + h * DH + lane, 0.f);
      }
    }
  }
}

size_t fwd_smem(int Lq, int Lk) {
  return sizeof(float) * ((size_t)2 * KC * KPAD + (size_t)Lq * KPAD + (size_t)NWARPS * 2 * KC) + sizeof(int) * (size_t)Lk;
}
size_t bwd_smem(int Lq, int Lk) {
  return sizeof(float) * ((size_t)2 * KCB * KPAD + (size_t)2 * Lq * KPAD + (size_t)2 * Lq * (KCB + 1)) +
         sizeof(int) * (size_t)Lk;
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 227 * 1024) {
    fira_set_error(FIRA_ERR_SHAPE, "%s: needs %zu B shared memory (> 227 KB)", name, bytes);
    return FIRA_ERR_SHAPE;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { fira_set_error(FIRA_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e)); return FIRA_ERR_CUDA; }
  }
  return FIRA_OK;
}

int check_layout(const char* name, const void* p, long ld, int dtype) {
  const long esz = dtype == FIRA_F32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0 || (ld * esz) % 16 != 0) {
    fira_set_error(FIRA_ERR_ALIGN, "%s: operands must be 16-B aligned with 16-B row pitch", name);
    return FIRA_ERR_ALIGN;
  }
  return FIRA_OK;
}

}  // namespace

namespace {

int attn_fwd_impl(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                  const unsigned char* key_mask, const int* ranges, long kv_rows, int max_chunks, int causal, void* ctx,
                  long ldo, float* stats, int B, int H, int Lq, int Lk, int d_head, int dtype, void* stream) {
  FIRA_CHECK_ARG(d_head == DH, FIRA_ERR_SHAPE, "attn_fwd: d_head %d != 32", d_head);
  FIRA_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0 && Lq <= LQ_MAX, FIRA_ERR_SHAPE, "attn_fwd: shape (Lq <= 32)");
  FIRA_CHECK_ARG(!causal || (Lq == Lk && !ranges), FIRA_ERR_SHAPE, "attn_fwd: causal needs Lq == Lk and no ranges");
  FIRA_CHECK_ARG(key_mask || ranges, FIRA_ERR_ARG, "attn_fwd: key_mask may only be NULL with ranges");
  FIRA_CHECK_ARG(dtype == FIRA_F32 || dtype == FIRA_BF16, FIRA_ERR_DTYPE, "attn_fwd: dtype %d", dtype);
  int rc;
  if ((rc = check_layout("attn_fwd", q, ldq, dtype)) || (rc = check_layout("attn_fwd", k, ldk, dtype)) ||
      (rc = check_layout("attn_fwd", v, ldv, dtype)))
    return rc;
  if (use_tc(dtype, B, H, Lq, Lk, d_head, ldk, ldv, max_chunks))
    return fira_attn_tc_fwd(q, ldq, k, ldk, v, ldv, key_mask, ranges, kv_rows, causal, ctx, ldo, stats, B, H, Lq, Lk, stream);
  AttnArgs a{q, ldq, k, ldk, v, ldv, key_mask, ranges, causal, B, H, Lq, Lk, 1.f / sqrtf((float)d_head)};
  const size_t smem = fwd_smem(Lq, Lk);
  if (dtype == FIRA_F32) {
    if ((rc = set_smem(attn_fwd_kernel<float>, smem, "attn_fwd"))) return rc;
    launch_k(attn_fwd_kernel<float>, dim3(B * H), dim3(NTHR), smem, (cudaStream_t)stream, a, (float*)ctx, ldo, stats);
  } else {
    if ((rc = set_smem(attn_fwd_kernel<__nv_bfloat16>, smem, "attn_fwd"))) return rc;
    launch_k(attn_fwd_kernel<__nv_bfloat16>, dim3(B * H), dim3(NTHR), smem, (cudaStream_t)stream, a, (__nv_bfloat16*)ctx, ldo, stats);
  }
  FIRA_CHECK_LAUNCH("fira_attn_fwd");
  return FIRA_OK;
}

int attn_bwd_impl(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                  const unsigned char* key_mask, const int* ranges, long kv_rows, int max_chunks, int causal,
                  const void* ctx, const void* d_ctx, long ldo, const float* stats, void* dq, long lddq, void* dk,
                  long lddk, void* dv, long lddv, int B, int H, int Lq, int Lk, int d_head, int dtype, void* stream) {
  FIRA_CHECK_ARG(d_head == DH, FIRA_ERR_SHAPE, "attn_bwd: d_head %d != 32", d_head);
  FIRA_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0 && Lq <= LQ_MAX, FIRA_ERR_SHAPE, "attn_bwd: shape (Lq <= 32)");
  FIRA_CHECK_ARG(key_mask || ranges, FIRA_ERR_ARG, "attn_bwd: key_mask may only be NULL with ranges");
  FIRA_CHECK_ARG(dtype == FIRA_F32 || dtype == FIRA_BF16, FIRA_ERR_DTYPE, "attn_bwd: dtype %d", dtype);
  int rc;
  if ((rc = check_layout("attn_bwd", q, ldq, dtype)) || (rc = check_layout("attn_bwd", k, ldk, dtype)) ||
      (rc = check_layout("attn_bwd", v, ldv, dtype)) || (rc = check_layout("attn_bwd", ctx, ldo, dtype)) ||
      (rc = check_layout("attn_bwd", d_ctx, ldo, dtype)))
    return rc;
  if (use_tc(dtype, B, H, Lq, Lk, d_head, ldk, ldv, max_chunks) && (lddk % 8) == 0 && (lddv % 8) == 0 &&
      (lddq % 8) == 0)
    return fira_attn_tc_bwd(q, ldq, k, ldk, v, ldv, key_mask, ranges, kv_rows, causal, ctx, d_ctx, ldo, stats, dq, lddq,
                            dk, lddk, dv, lddv, B, H, Lq, Lk, stream);
  AttnArgs a{q, ldq, k, ldk, v, ldv, key_mask, ranges, causal, B, H, Lq, Lk, 1.f / sqrtf((float)d_head)};
  const size_t smem = bwd_smem(Lq, Lk);
  if (dtype == FIRA_F32) {
    if ((rc = set_smem(attn_bwd_kernel<float>, smem, "attn_bwd"))) return rc;
    launch_k(attn_bwd_kernel<float>, dim3(B * H), dim3(NTHR), smem, (cudaStream_t)stream, 
        a, (const float*)ctx, (const float*)d_ctx, ldo, stats, (float*)dq, lddq, (float*)dk, lddk, (float*)dv, lddv);
  } else {
    if ((rc = set_smem(attn_bwd_kernel<__nv_bfloat16>, smem, "attn_bwd"))) return rc;
    launch_k(attn_bwd_kernel<__nv_bfloat16>, dim3(B * H), dim3(NTHR), smem, (cudaStream_t)stream, 
        a, (const __nv_bfloat16*)ctx, (const __nv_bfloat16*)d_ctx, ldo, stats, (__nv_bfloat16*)dq, lddq,
        (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv);
  }
  FIRA_CHECK_LAUNCH("fira_attn_bwd");
  return FIRA_OK;
}

}  // namespace

extern "C" {

int fira_attn_fwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                  const unsigned char* key_mask, int causal, void* ctx, long ldo, float* stats, int B, int H, int Lq,
                  int Lk, int d_head, int dtype, void* stream) {
  FIRA_CHECK_ARG(key_mask, FIRA_ERR_ARG, "attn_fwd: null key_mask");
  return attn_fwd_impl(q, ldq, k, ldk, v, ldv, key_mask, nullptr, (long)B * Lk, (Lk + 127) / 128, causal, ctx, ldo, stats,
                       B, H, Lq, Lk, d_head, dtype, stream);
}

int fira_attn_bwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                  const unsigned char* key_mask, int causal, const void* ctx, const void* d_ctx, long ldo,
                  const float* stats, void* dq, long lddq, void* dk, long lddk, void* dv, long lddv, int B, int H,
                  int Lq, int Lk, int d_head, int dtype, void* stream) {
  FIRA_CHECK_ARG(key_mask, FIRA_ERR_ARG, "attn_bwd: null key_mask");
  return attn_bwd_impl(q, ldq, k, ldk, v, ldv, key_mask, nullptr, (long)B * Lk, (Lk + 127) / 128, causal, ctx, d_ctx, ldo,
                       stats, dq, lddq, dk, lddk, dv, lddv, B, H, Lq, Lk, d_head, dtype, stream);
}

int fira_attn_packed_fwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv, const int* ranges,
                         long kv_rows, const unsigned char* key_mask, int mask_pitch, int max_chunks, void* ctx, long ldo,
                         float* stats, int B, int H, int Lq, int d_head, int dtype, void* stream) {
  FIRA_CHECK_ARG(ranges && kv_rows > 0 && max_chunks > 0, FIRA_ERR_ARG, "attn_packed_fwd: ranges / kv_rows / max_chunks");
  return attn_fwd_impl(q, ldq, k, ldk, v, ldv, key_mask, ranges, kv_rows, max_chunks, 0, ctx, ldo, stats, B, H, Lq,
                       mask_pitch, d_head, dtype, stream);
}

int fira_attn_packed_bwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv, const int* ranges,
                         long kv_rows, const unsigned char* key_mask, int mask_pitch, int max_chunks, const void* ctx,
                         const void* d_ctx, long ldo, const float* stats, void* dq, long lddq, void* dk, long lddk,
                         void* dv, long lddv, int B, int H, int Lq, int d_head, int dtype, void* stream) {
  FIRA_CHECK_ARG(ranges && kv_rows > 0 && max_chunks > 0, FIRA_ERR_ARG, "attn_packed_bwd: ranges / kv_rows / max_chunks");
  return attn_bwd_impl(q, ldq, k, ldk, v, ldv, key_mask, ranges, kv_rows, max_chunks, 0, ctx, d_ctx, ldo, stats, dq, lddq,
                       dk, lddk, dv, lddv, B, H, Lq, mask_pitch, d_head, dtype, stream);
}

}  // extern "C"
