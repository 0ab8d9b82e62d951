// Multi-head attention core for the parity (fp32) path: one CTA per (batch, head) keeps K, V and the
// 30 query rows of that head on chip (the whole problem fits in one SM's shared memory: Lq = 30,
// Lk <= 370, d_head = 32).  gnn_transformer.py:144-156:
//     S = Q K^T / sqrt(d_head);  S[mask == 0] = -1e9;  P = softmax(S);  ctx = P V
// mask = key padding (and causal for decoder self-attention, gnn_transformer.py:117).  No dropout on P.
// Forward saves (row max, row sum) so backward recomputes P exactly instead of storing B*8*30*370 floats.
// The throughput path (attention_tc.cu) does the two contractions on tcgen05 instead.
#include "common.cuh"
#include "fira_b200.h"

namespace {

constexpr int DH = 32;           // head dim (256 / 8)
constexpr int KPAD = DH + 1;     // conflict-free column reads of K/V tiles
constexpr int NWARPS = 8;

struct AttnArgs {
  const void* q; long ldq;       // row (b*Lq + t), head h at column h*32
  const void* k; long ldk;       // row (b*Lk + s)
  const void* v; long ldv;
  const unsigned char* key_mask; // [B, Lk], 1 = attend
  int causal;
  int B, H, Lq, Lk;
  float scale;
};

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, const T* src, long ld, int rows, int tid, int nthr) {
  // rows x 32 -> dst[row][KPAD].  16-byte loads (a head slice of a row is 128 B fp32 / 64 B bf16), several
  // independent loads in flight per thread: the scalar version of round 1 was latency-bound (one 4-byte
  // load per thread per trip, ~40 us per 370-row tile).
  constexpr int EPV = 16 / sizeof(T);          // elements per 16-byte vector
  constexpr int VPR = DH / EPV;                // vectors per row
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((ld * sizeof(T)) % 16 == 0);
  if (vec_ok) {
#pragma unroll 4
    for (int idx = tid; idx < rows * VPR; idx += nthr) {
      const int r = idx / VPR, c = idx % VPR;
      const uint4 raw = *reinterpret_cast<const uint4*>(src + (long)r * ld + c * EPV);
      float* o = dst + r * KPAD + c * EPV;
      if constexpr (sizeof(T) == 4) {
        o[0] = __uint_as_float(raw.x); o[1] = __uint_as_float(raw.y);
        o[2] = __uint_as_float(raw.z); o[3] = __uint_as_float(raw.w);
      } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); o[2 * q] = f.x; o[2 * q + 1] = f.y; }
      }
    }
  } else {
    for (int idx = tid; idx < rows * DH; idx += nthr) {
      int r = idx / DH, d = idx % DH;
      dst[r * KPAD + d] = Act<T>::ld(src + (long)r * ld + d);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(NWARPS * 32) attn_fwd_kernel(AttnArgs a, T* __restrict__ ctx, long ldo,
                                                               float* __restrict__ stats /* [B,H,Lq,2] */) {
  extern __shared__ float smem[];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* Ks = smem;                         // [Lk][KPAD]
  float* Vs = Ks + a.Lk * KPAD;             // [Lk][KPAD]
  float* Qs = Vs + a.Lk * KPAD;             // [Lq][KPAD]
  float* Ps = Qs + a.Lq * KPAD;             // [NWARPS][Lk]
  load_tile(Ks, (const T*)a.k + (long)b * a.Lk * a.ldk + h * DH, a.ldk, a.Lk, threadIdx.x, blockDim.x);
  load_tile(Vs, (const T*)a.v + (long)b * a.Lk * a.ldv + h * DH, a.ldv, a.Lk, threadIdx.x, blockDim.x);
  load_tile(Qs, (const T*)a.q + (long)b * a.Lq * a.ldq + h * DH, a.ldq, a.Lq, threadIdx.x, blockDim.x);
  __syncthreads();
  const unsigned char* km = a.key_mask + (long)b * a.Lk;
  float* P = Ps + warp * a.Lk;
  for (int t = warp; t < a.Lq; t += NWARPS) {
    float qreg[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) qreg[d] = Qs[t * KPAD + d];
    float mx = -INFINITY;
    for (int s = lane; s < a.Lk; s += 32) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc = fmaf(qreg[d], Ks[s * KPAD + d], acc);
      acc *= a.scale;
      const bool ok = km[s] && (!a.causal || s <= t);
      acc = ok ? acc : kMaskFill;
      P[s] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < a.Lk; s += 32) { float e = expf(P[s] - mx); P[s] = e; sum += e; }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    float o = 0.f;
    for (int s = 0; s < a.Lk; ++s) o = fmaf(P[s], Vs[s * KPAD + lane], o);
    Act<T>::st(ctx + ((long)b * a.Lq + t) * ldo + h * DH + lane, o * inv);
    if (lane == 0 && stats) {
      float* st = stats + (((long)b * a.H + h) * a.Lq + t) * 2;
      st[0] = mx; st[1] = sum;
    }
    __syncwarp();
  }
}

// Backward: dV = P^T dO, dP = dO V^T, dS = P * (dP - rowsum(P*dP)), dQ = scale dS K, dK = scale dS^T Q.
template <typename T>
__global__ void __launch_bounds__(NWARPS * 32) attn_bwd_kernel(AttnArgs a, const T* __restrict__ d_ctx, long ldo,
                                                               const float* __restrict__ stats, T* __restrict__ dq,
                                                               long lddq, T* __restrict__ dk, long lddk,
                                                               T* __restrict__ dv, long lddv) {
  extern __shared__ float smem[];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int LkP = a.Lk + 1;
  float* Ks = smem;                         // [Lk][KPAD]
  float* Vs = Ks + a.Lk * KPAD;             // [Lk][KPAD]
  float* Qs = Vs + a.Lk * KPAD;             // [Lq][KPAD]
  float* Os = Qs + a.Lq * KPAD;             // [Lq][KPAD]   dO
  float* Pm = Os + a.Lq * KPAD;             // [Lq][LkP]    P
  float* Sm = Pm + a.Lq * LkP;              // [Lq][LkP]    dS
  load_tile(Ks, (const T*)a.k + (long)b * a.Lk * a.ldk + h * DH, a.ldk, a.Lk, threadIdx.x, blockDim.x);
  load_tile(Vs, (const T*)a.v + (long)b * a.Lk * a.ldv + h * DH, a.ldv, a.Lk, threadIdx.x, blockDim.x);
  load_tile(Qs, (const T*)a.q + (long)b * a.Lq * a.ldq + h * DH, a.ldq, a.Lq, threadIdx.x, blockDim.x);
  load_tile(Os, d_ctx + (long)b * a.Lq * ldo + h * DH, ldo, a.Lq, threadIdx.x, blockDim.x);
  __syncthreads();
  const unsigned char* km = a.key_mask + (long)b * a.Lk;
  for (int t = warp; t < a.Lq; t += NWARPS) {
    float qreg[DH], oreg[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { qreg[d] = Qs[t * KPAD + d]; oreg[d] = Os[t * KPAD + d]; }
    const float* st = stats + (((long)b * a.H + h) * a.Lq + t) * 2;
    const float mx = st[0], inv = 1.f / st[1];
    float delta = 0.f;
    for (int s = lane; s < a.Lk; s += 32) {
      float acc = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        acc = fmaf(qreg[d], Ks[s * KPAD + d], acc);
        dp = fmaf(oreg[d], Vs[s * KPAD + d], dp);
      }
      acc *= a.scale;
      const bool ok = km[s] && (!a.causal || s <= t);
      acc = ok ? acc : kMaskFill;
      const float p = expf(acc - mx) * inv;
      Pm[t * LkP + s] = p;
      Sm[t * LkP + s] = dp;
      delta = fmaf(p, dp, delta);
    }
    delta = warp_sum(delta);
    // masked_fill overwrote the masked scores: no gradient reaches them even when P != 0 there
    // (a fully masked row has uniform P)
    for (int s = lane; s < a.Lk; s += 32) {
      const bool ok = km[s] && (!a.causal || s <= t);
      Sm[t * LkP + s] = ok ? Pm[t * LkP + s] * (Sm[t * LkP + s] - delta) : 0.f;
    }
    __syncwarp();
    float g = 0.f;   // dQ[t][lane]
    for (int s = 0; s < a.Lk; ++s) g = fmaf(Sm[t * LkP + s], Ks[s * KPAD + lane], g);
    Act<T>::st(dq + ((long)b * a.Lq + t) * lddq + h * DH + lane, g * a.scale);
  }
  __syncthreads();
  for (int s = warp; s < a.Lk; s += NWARPS) {
    float gk = 0.f, gv = 0.f;
    for (int t = 0; t < a.Lq; ++t) {
      gk = fmaf(Sm[t * LkP + s], Qs[t * KPAD + lane], gk);
      gv = fmaf(Pm[t * LkP + s], Os[t * KPAD + lane], gv);
    }
    Act<T>::st(dk + ((long)b * a.Lk + s) * lddk + h * DH + lane, gk * a.scale);
    Act<T>::st(dv + ((long)b * a.Lk + s) * lddv + h * DH + lane, gv);
  }
}

size_t fwd_smem(int Lq, int Lk) { return sizeof(float) * ((size_t)2 * Lk * KPAD + (size_t)Lq * KPAD + (size_t)NWARPS * Lk); }
size_t bwd_smem(int Lq, int Lk) {
  return sizeof(float) * ((size_t)2 * Lk * KPAD + (size_t)2 * Lq * KPAD + (size_t)2 * Lq * (Lk + 1));
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

}  // namespace

extern "C" {

int fira_attn_fwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                  const unsigned char* key_mask, int causal, void* ctx, long ldo, float* stats, int B, int H, int Lq,
                  int Lk, int d_head, int dtype, void* stream) {
  FIRA_CHECK_ARG(d_head == DH, FIRA_ERR_SHAPE, "attn_fwd: d_head %d != 32", d_head);
  FIRA_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0, FIRA_ERR_SHAPE, "attn_fwd: shape");
  FIRA_CHECK_ARG(!causal || Lq == Lk, FIRA_ERR_SHAPE, "attn_fwd: causal needs Lq == Lk");
  AttnArgs a{q, ldq, k, ldk, v, ldv, key_mask, causal, B, H, Lq, Lk, 1.f / sqrtf((float)d_head)};
  const size_t smem = fwd_smem(Lq, Lk);
  int rc;
  if (dtype == FIRA_F32) {
    if ((rc = set_smem(attn_fwd_kernel<float>, smem, "attn_fwd"))) return rc;
    attn_fwd_kernel<float><<<B * H, NWARPS * 32, smem, (cudaStream_t)stream>>>(a, (float*)ctx, ldo, stats);
  } else if (dtype == FIRA_BF16) {
    if ((rc = set_smem(attn_fwd_kernel<__nv_bfloat16>, smem, "attn_fwd"))) return rc;
    attn_fwd_kernel<__nv_bfloat16><<<B * H, NWARPS * 32, smem, (cudaStream_t)stream>>>(a, (__nv_bfloat16*)ctx, ldo, stats);
  } else { fira_set_error(FIRA_ERR_DTYPE, "attn_fwd: dtype %d", dtype); return FIRA_ERR_DTYPE; }
  FIRA_CHECK_LAUNCH("fira_attn_fwd");
  return FIRA_OK;
}

int fira_attn_bwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                  const unsigned char* key_mask, int causal, const void* d_ctx, long ldo, const float* stats, void* dq,
                  long lddq, void* dk, long lddk, void* dv, long lddv, int B, int H, int Lq, int Lk, int d_head,
                  int dtype, void* stream) {
  FIRA_CHECK_ARG(d_head == DH, FIRA_ERR_SHAPE, "attn_bwd: d_head %d != 32", d_head);
  FIRA_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0, FIRA_ERR_SHAPE, "attn_bwd: shape");
  AttnArgs a{q, ldq, k, ldk, v, ldv, key_mask, causal, B, H, Lq, Lk, 1.f / sqrtf((float)d_head)};
  const size_t smem = bwd_smem(Lq, Lk);
  int rc;
  if (dtype == FIRA_F32) {
    if ((rc = set_smem(attn_bwd_kernel<float>, smem, "attn_bwd"))) return rc;
    attn_bwd_kernel<float><<<B * H, NWARPS * 32, smem, (cudaStream_t)stream>>>(
        a, (const float*)d_ctx, ldo, stats, (float*)dq, lddq, (float*)dk, lddk, (float*)dv, lddv);
  } else if (dtype == FIRA_BF16) {
    if ((rc = set_smem(attn_bwd_kernel<__nv_bfloat16>, smem, "attn_bwd"))) return rc;
    attn_bwd_kernel<__nv_bfloat16><<<B * H, NWARPS * 32, smem, (cudaStream_t)stream>>>(
        a, (const __nv_bfloat16*)d_ctx, ldo, stats, (__nv_bfloat16*)dq, lddq, (__nv_bfloat16*)dk, lddk,
        (__nv_bfloat16*)dv, lddv);
  } else { fira_set_error(FIRA_ERR_DTYPE, "attn_bwd: dtype %d", dtype); return FIRA_ERR_DTYPE; }
  FIRA_CHECK_LAUNCH("fira_attn_bwd");
  return FIRA_OK;
}

}  // extern "C"
