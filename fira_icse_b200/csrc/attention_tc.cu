// Multi-head attention core on the 5th-generation tensor cores (bf16 throughput mode), gnn_transformer.py:144-156:
//     S = Q K^T / sqrt(32);  S[mask == 0] = -1e9;  P = softmax(S);  ctx = P V          (+ the backward of exactly that)
//
// The per-head problem (30 x Lk x 32) is far below a tcgen05.mma tile (M = 128), and every head has its own K/V slice,
// so the kernel works on a HEAD GROUP: one CTA per (commit, 4 heads).  The 128 rows of the UMMA tile are
// (head hl, query t) = hl*32 + t, and the A operand is the query block REPLICATED and MASKED per head:
//     A'[(hl,t), f] = Q[t, 128 g + f]  if feature f belongs to head hl (f / 32 == hl), else 0          (128 x 128)
// With that operand one M128 x N128 x K128 product against the K rows AS THEY LIE IN MEMORY (K-major, 128 features of
// the group) yields all four heads' score rows at once; the zeros make the cross-head terms vanish, the tensor core
// does 4x redundant MACs on a pipe that is otherwise idle.  The same trick runs backwards:
//     O'  = P V          (M = (hl,t), K = keys, N = 128 features; row (hl,t) keeps its own head's 32 columns)
//     dP' = dO' V^T ,  dQ' = dS K ,  dK = dS^T A'(Q) ,  dV = P^T A'(dO)      (A' as MN-major B: cross-head terms are 0)
// K and V tiles arrive by TMA (SWIZZLE_128B); one [128 keys x 64 features] box serves as K-major B (scores) and as
// MN-major B (P V, dS K) -- the bytes are the same, only the descriptor differs.  P and dS are written by the softmax
// threads (thread = row) into the same swizzled layout, where they serve as K-major A (P V, dS K) and as MN-major A
// (P^T dO, dS^T Q).  Accumulators live in TMEM (512 columns); softmax statistics are thread-local (lane = row).
// Keys are processed in chunks of 128; forward keeps all score chunks in TMEM (Lk <= 384) and makes two passes
// (max, then exp / sum / P), so no online rescaling is needed.
//
// Semantics kept from the FFMA kernels (attention.cu): scale before mask, -1e9 fill (a fully masked row is uniform
// over all Lk keys), masked keys get exactly zero dK / dV unless the whole row is masked, statistics = (row max, row
// sum) [B,H,Lq,2].
#include "tc_common.cuh"
#include "fira_b200.h"

namespace attn_tc {

using namespace tc;

constexpr int DH = 32, HG = 4, GF = HG * DH;         // head dim, heads per group, features per group (128)
constexpr int KC = 128;                              // keys per chunk
constexpr int MAX_CH = 3;                            // Lk <= 384
constexpr int TILE = 128 * 128 * 2;                  // one [128 rows x 128 cols] bf16 tile = two 16 KB panels
constexpr int PANEL = 16384;
constexpr int THREADS = 192;                         // warps 0-3: softmax / epilogue (TMEM lane quarters), 4: TMA + MMA, 5: helper

struct Args {
  const __nv_bfloat16* q; long ldq;
  const unsigned char* key_mask;                     // [B, Lk] (NULL with ranges: every key of the ranges is valid)
  const int* ranges;                                 // NULL, or [B][4] = {first row, rows, first row, rows}: the keys of
                                                     // commit b are two row ranges of k / v (packed batches); Lk = mask pitch
  int causal, B, H, Lq, Lk;
  float scale;
  __nv_bfloat16* ctx; long ldo;                      // fwd out / bwd: forward output
  float* stats;
  // backward
  const __nv_bfloat16* d_ctx;
  __nv_bfloat16* dq; long lddq;
  __nv_bfloat16* dk; long lddk;
  __nv_bfloat16* dv; long lddv;
};

__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }

// A'(X)[(hl,t), f] tile for head group g from a [B*Lq, ld] matrix: zero-filled, then head hl's 32 features of row t.
// tile layout: 2 panels (64 features) x [128 rows x 128 B], SWIZZLE_128B.
__device__ __forceinline__ void build_masked_tile(unsigned char* tile, const __nv_bfloat16* x, long ld, int b, int g, int Lq,
                                                  int tid, int nthr) {
  for (int i = tid; i < TILE / 16; i += nthr) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // 128 rows x 4 chunks of 16 B
  for (int i = tid; i < 128 * 4; i += nthr) {
    const int m = i >> 2, c = i & 3;
    const int hl = m >> 5, t = m & 31;
    if (t < Lq) {
      const uint4 v = ldg16(x + ((long)b * Lq + t) * ld + g * GF + hl * DH + c * 8);
      const int f = hl * DH + c * 8;                 // feature within the group
      *reinterpret_cast<uint4*>(tile + (f >> 6) * PANEL + sw128_offset(m, (f & 63) >> 3)) = v;
    }
  }
}

// Key chunks of one commit: chunk c = rows [row[c], row[c] + n[c]) of k / v, its keys are mask positions moff[c] + i.
struct Chunks { int nch; int row[MAX_CH]; int n[MAX_CH]; int moff[MAX_CH]; };

__device__ __forceinline__ void make_chunks(Chunks& ch, const int* ranges, int b, int Lk) {
  int n = 0;
  if (ranges) {
    const int s0 = ranges[4 * b], l0 = ranges[4 * b + 1], s1 = ranges[4 * b + 2], l1 = ranges[4 * b + 3];
    for (int o = 0; o < l0 && n < MAX_CH; o += KC, ++n) { ch.row[n] = s0 + o; ch.n[n] = min(KC, l0 - o); ch.moff[n] = o; }
    for (int o = 0; o < l1 && n < MAX_CH; o += KC, ++n) { ch.row[n] = s1 + o; ch.n[n] = min(KC, l1 - o); ch.moff[n] = l0 + o; }
  } else {
    for (int o = 0; o < Lk && n < MAX_CH; o += KC, ++n) { ch.row[n] = b * Lk + o; ch.n[n] = min(KC, Lk - o); ch.moff[n] = o; }
  }
  ch.nch = n;
}

// Key validity as bit masks, one word per 32-key group of a chunk:
//   exist[c*4+j] bit i : key j*32+i of chunk c is a key of this commit (below the chunk's key count)
//   bits [c*4+j] bit i : ... and its mask byte is set
// A chunk without a single valid key is DROPPED from the table when the commit has valid keys elsewhere (its
// probabilities and gradients are exactly zero: exp(-1e9 - max) == 0 in fp32); `rm` keeps the dropped chunks so that
// the backward can write their zero dK / dV rows.  A commit / row without any valid key keeps everything: every score
// is -1e9 there and the softmax is uniform over ALL keys, like the reference's masked_fill + softmax.
struct KeyBits { uint32_t bits[MAX_CH * 4]; uint32_t exist[MAX_CH * 4]; };

__device__ __forceinline__ void prepare_keys(Chunks& ch, Chunks& rm, KeyBits& kb, const unsigned char* s_mask, int causal,
                                             int warp, int lane) {
  if (warp == 0) {
    unsigned has = 0;
    for (int c = 0; c < ch.nch; ++c) {
      bool v = false;
      for (int i = lane; i < ch.n[c]; i += 32) v |= s_mask[ch.moff[c] + i] != 0;
      if (__any_sync(0xffffffffu, v)) has |= 1u << c;
    }
    if (lane == 0) {
      int k = 0, r = 0;
      const bool drop = has != 0 && !causal;
      for (int c = 0; c < ch.nch; ++c) {
        if (!drop || ((has >> c) & 1)) { ch.row[k] = ch.row[c]; ch.n[k] = ch.n[c]; ch.moff[k] = ch.moff[c]; ++k; }
        else { rm.row[r] = ch.row[c]; rm.n[r] = ch.n[c]; rm.moff[r] = ch.moff[c]; ++r; }
      }
      ch.nch = k; rm.nch = r;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < ch.nch * KC; idx += THREADS) {       // THREADS % 32 == 0: a warp covers one group
    const int c = idx >> 7, i = idx & 127;
    const bool ex = i < ch.n[c];
    const bool ok = ex && s_mask[ch.moff[c] + i] != 0;
    const unsigned bv = __ballot_sync(0xffffffffu, ok), be = __ballot_sync(0xffffffffu, ex);
    if (lane == 0) { kb.bits[idx >> 5] = bv; kb.exist[idx >> 5] = be; }
  }
  __syncthreads();
}

// valid keys of group j of chunk c for query row t
__device__ __forceinline__ uint32_t row_bits(const KeyBits& kb, const Chunks& ch, int c, int j, int t, int causal) {
  uint32_t v = kb.bits[c * 4 + j];
  if (causal) {                                     // key position <= t (causal attention has one chunk, moff = 0)
    const int lo = ch.moff[c] + j * 32;
    v &= t < lo ? 0u : (t - lo >= 31 ? 0xffffffffu : ((2u << (t - lo)) - 1u));
  }
  return v;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;

// =================================================================================================== forward
// smem: A'(Q) 32 KB | KV[3] 96 KB | P[2] 64 KB
__global__ void __launch_bounds__(THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, Args a) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long kv_full[MAX_CH], v_full[MAX_CH], s_full, p_full[2], p_empty[2], o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned char s_mask[MAX_CH * KC];
  __shared__ Chunks ch, rm;
  __shared__ KeyBits kb;
  const uint32_t base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - smem_addr(smem_raw));
  constexpr uint32_t OFF_AQ = 0, OFF_KV = TILE, OFF_P = OFF_KV + MAX_CH * TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x >> 1, g = blockIdx.x & 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_CH; ++i) { mbar_init(smem_addr(&kv_full[i]), 1); mbar_init(smem_addr(&v_full[i]), 1); }
    mbar_init(smem_addr(&s_full), 1);
    mbar_init(smem_addr(&o_full), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_addr(&p_full[i]), 128); mbar_init(smem_addr(&p_empty[i]), 1); }
    mbar_init_fence();
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 4) tmem_alloc(smem_addr(&tmem_slot), 512);
  pdl_wait(); pdl_trigger();       // PDL: the prologue above overlapped the previous kernel's tail (common.cuh)
  if (threadIdx.x == 0) make_chunks(ch, a.ranges, b, a.Lk);
  for (int i = threadIdx.x; i < MAX_CH * KC; i += THREADS)
    s_mask[i] = i < a.Lk ? (a.key_mask ? a.key_mask[(long)b * a.Lk + i] : (unsigned char)1) : (unsigned char)0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  prepare_keys(ch, rm, kb, s_mask, a.causal, warp, lane);
  const int nch = ch.nch;
  if (warp == 4 && lane == 0) {                      // K chunks can fly while the A' tile is built
    for (int c = 0; c < nch; ++c) {
      const uint32_t dst = base + OFF_KV + c * TILE, bar = smem_addr(&kv_full[c]);
      mbar_expect_tx(bar, TILE);
      tma_load_2d(dst, &tmK, g * GF, ch.row[c], bar);
      tma_load_2d(dst + PANEL, &tmK, g * GF + 64, ch.row[c], bar);
    }
  }
  build_masked_tile(sm + OFF_AQ, a.q, a.ldq, b, g, a.Lq, threadIdx.x, THREADS);
  fence_proxy_async();
  __syncthreads();

  if (warp == 4) {
    if (lane == 0) {
      // ---- S chunks = A'(Q) K^T : K-major A (2 k-blocks of 64 features), K-major B (key rows)
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KC, false, false);
      for (int c = 0; c < nch; ++c) {
        mbar_wait(smem_addr(&kv_full[c]), 0);
        tc_fence_after();
#pragma unroll
        for (int kb_ = 0; kb_ < 2; ++kb_)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + c * KC, make_desc(base + OFF_AQ + kb_ * PANEL + k * 32, 16, 1024),
                      make_desc(base + OFF_KV + c * TILE + kb_ * PANEL + k * 32, 16, 1024), idesc_s, (kb_ | k) ? 1u : 0u);
      }
      umma_commit(smem_addr(&s_full));
      mbar_wait(smem_addr(&s_full), 0);               // the K tiles have been read: reuse their buffers for V
      for (int c = 0; c < nch; ++c) {
        const uint32_t dst = base + OFF_KV + c * TILE, bar = smem_addr(&v_full[c]);
        mbar_expect_tx(bar, TILE);
        tma_load_2d(dst, &tmV, g * GF, ch.row[c], bar);
        tma_load_2d(dst + PANEL, &tmV, g * GF + 64, ch.row[c], bar);
      }
      // ---- O' += P_c V_c : K-major A (P, 2 k-blocks of 64 keys), MN-major B (V: K = key rows, N = 128 features)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, GF, false, true);
      for (int c = 0; c < nch; ++c) {
        mbar_wait(smem_addr(&v_full[c]), 0);
        mbar_wait(smem_addr(&p_full[c & 1]), (c >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)                    // 8 steps of 16 keys
          umma_bf16(tmem + MAX_CH * KC,
                    make_desc(base + OFF_P + (c & 1) * TILE + (k >> 2) * PANEL + (k & 3) * 32, 16, 1024),
                    make_desc(base + OFF_KV + c * TILE + k * 2048, PANEL, 1024), idesc_o, (c | k) ? 1u : 0u);
        umma_commit(smem_addr(&p_empty[c & 1]));
      }
      umma_commit(smem_addr(&o_full));
    }
  } else if (warp < 4) {
    // ---- softmax + epilogue: thread = row (hl = warp, t = lane)
    const int t = lane;
    const bool live = t < a.Lq;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    // does this row see any valid key at all?  (no: every score is -1e9 -> uniform over all existing keys)
    uint32_t any = 0;
    for (int c = 0; c < nch; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) any |= row_bits(kb, ch, c, j, t, a.causal);
    const bool rowfilled = any == 0;
    if (lane == 0) mbar_wait(smem_addr(&s_full), 0);
    __syncwarp();
    tc_fence_after();
    // pass A: row maximum of the raw scores over the valid keys (the scale is positive)
    float mraw = -INFINITY;
    for (int c = 0; c < nch; ++c)
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const uint32_t vb = row_bits(kb, ch, c, j, t, a.causal);
        if (__any_sync(0xffffffffu, vb != 0)) {       // warp-uniform: tcgen05.ld is .sync.aligned
          uint32_t r[32];
          tmem_ld32(trow + c * KC + j * 32, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) if ((vb >> i) & 1) mraw = fmaxf(mraw, __uint_as_float(r[i]));
        }
      }
    const float mx = rowfilled ? kMaskFill : mraw * a.scale;            // what the reference's softmax subtracts
    const float k2 = a.scale * kLog2e, m2 = mraw * k2;
    float sum = 0.f;
    const int m = warp * 32 + lane;
    for (int c = 0; c < nch; ++c) {
      if (c >= 2) { if (lane == 0) mbar_wait(smem_addr(&p_empty[c & 1]), ((c >> 1) - 1) & 1); __syncwarp(); }
      unsigned char* ptile = sm + OFF_P + (c & 1) * TILE;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const uint32_t vb = row_bits(kb, ch, c, j, t, a.causal), eb = kb.exist[c * 4 + j];
        float e[32];
        if (__any_sync(0xffffffffu, vb != 0 && !rowfilled)) {
          uint32_t r[32];
          tmem_ld32(trow + c * KC + j * 32, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // the MMA consumes bf16(e): sum the ROUNDED values so that P rows are normalised exactly
            const float x = rowfilled ? (((eb >> i) & 1) ? 1.f : 0.f)
                                      : (((vb >> i) & 1) ? fast_exp2(fmaf(__uint_as_float(r[i]), k2, -m2)) : 0.f);
            e[i] = live ? __bfloat162float(__float2bfloat16_rn(x)) : 0.f;
            sum += e[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) { e[i] = (live && rowfilled && ((eb >> i) & 1)) ? 1.f : 0.f; sum += e[i]; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(e[q * 8 + 2 * i], e[q * 8 + 2 * i + 1]);
          const int key = j * 32 + q * 8;             // key within the chunk
          *reinterpret_cast<uint4*>(ptile + (key >> 6) * PANEL + sw128_offset(m, (key & 63) >> 3)) = v;
        }
      }
      fence_proxy_async();
      mbar_arrive(smem_addr(&p_full[c & 1]));
    }
    if (lane == 0) mbar_wait(smem_addr(&o_full), 0);
    __syncwarp();
    tc_fence_after();
    uint32_t r[32];
    tmem_ld32(trow + MAX_CH * KC + warp * DH, r);      // own head's 32 output columns
    if (live) {
      const float inv = 1.f / sum;
      __nv_bfloat16* dst = a.ctx + ((long)b * a.Lq + t) * a.ldo + g * GF + warp * DH;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(r[q * 8 + i]) * inv;
        Act<__nv_bfloat16>::store8(dst + q * 8, o);
      }
      if (a.stats) {
        float* st = a.stats + (((long)b * a.H + g * HG + warp) * a.Lq + t) * 2;
        st[0] = mx; st[1] = sum;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// =================================================================================================== backward
// smem: A'(Q) | A'(dO) | K | V | P | dS  (6 x 32 KB).  TMEM: S [0,128) dP [128,256) dQ [256,384) dK [384,512), dV reuses [0,128).
__global__ void __launch_bounds__(THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, Args a) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long kv_full, s_full, ds_full, g_full, epi_done;
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned char s_mask[MAX_CH * KC];
  __shared__ Chunks ch, rm;
  __shared__ KeyBits kb;
  const uint32_t base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - smem_addr(smem_raw));
  constexpr uint32_t OFF_AQ = 0, OFF_ADO = TILE, OFF_K = 2 * TILE, OFF_V = 3 * TILE, OFF_P = 4 * TILE, OFF_DS = 5 * TILE;
  constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256, T_DK = 384, T_DV = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x >> 1, g = blockIdx.x & 1;

  if (threadIdx.x == 0) {
    rm.nch = 0;
    mbar_init(smem_addr(&kv_full), 1);
    mbar_init(smem_addr(&s_full), 1);
    mbar_init(smem_addr(&ds_full), 128);
    mbar_init(smem_addr(&g_full), 1);
    mbar_init(smem_addr(&epi_done), 128);
    mbar_init_fence();
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 4) tmem_alloc(smem_addr(&tmem_slot), 512);
  pdl_wait(); pdl_trigger();       // PDL: the prologue above overlapped the previous kernel's tail (common.cuh)
  if (threadIdx.x == 0) make_chunks(ch, a.ranges, b, a.Lk);
  for (int i = threadIdx.x; i < MAX_CH * KC; i += THREADS)
    s_mask[i] = i < a.Lk ? (a.key_mask ? a.key_mask[(long)b * a.Lk + i] : (unsigned char)1) : (unsigned char)0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  prepare_keys(ch, rm, kb, s_mask, a.causal, warp, lane);
  const int nch = ch.nch;
  if (warp == 4 && lane == 0 && nch > 0) {           // the first K / V chunk flies while the A' tiles are built
    const uint32_t bar = smem_addr(&kv_full);
    mbar_expect_tx(bar, 2 * TILE);
    tma_load_2d(base + OFF_K, &tmK, g * GF, ch.row[0], bar);
    tma_load_2d(base + OFF_K + PANEL, &tmK, g * GF + 64, ch.row[0], bar);
    tma_load_2d(base + OFF_V, &tmV, g * GF, ch.row[0], bar);
    tma_load_2d(base + OFF_V + PANEL, &tmV, g * GF + 64, ch.row[0], bar);
  }
  // dropped chunks (no valid key, the commit has valid keys elsewhere): their dK / dV rows are exactly zero
  for (int c = 0; c < rm.nch; ++c)
    for (int i = threadIdx.x; i < rm.n[c] * (GF / 8); i += THREADS) {
      const int r = i / (GF / 8), q = i % (GF / 8);
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(a.dk + (long)(rm.row[c] + r) * a.lddk + g * GF + q * 8) = z;
      *reinterpret_cast<uint4*>(a.dv + (long)(rm.row[c] + r) * a.lddv + g * GF + q * 8) = z;
    }
  build_masked_tile(sm + OFF_AQ, a.q, a.ldq, b, g, a.Lq, threadIdx.x, THREADS);
  __syncthreads();
  build_masked_tile(sm + OFF_ADO, a.d_ctx, a.ldo, b, g, a.Lq, threadIdx.x, THREADS);
  fence_proxy_async();
  __syncthreads();

  if (warp == 4) {
    if (lane == 0) {
      constexpr uint32_t idesc_kk = make_idesc_bf16(128, KC, false, false);      // S, dP : K-major A, K-major B
      constexpr uint32_t idesc_kn = make_idesc_bf16(128, GF, false, true);       // dQ     : K-major A (dS), MN-major B (K)
      constexpr uint32_t idesc_nn = make_idesc_bf16(128, GF, true, true);        // dK, dV : MN-major A (dS^T / P^T), MN-major B (A')
      for (int c = 0; c < nch; ++c) {
        const uint32_t bar = smem_addr(&kv_full);
        if (c >= 1) {
          mbar_wait(smem_addr(&epi_done), (c - 1) & 1); // dK/dV of the previous chunk read out; K/V/P/dS free
          mbar_expect_tx(bar, 2 * TILE);
          tma_load_2d(base + OFF_K, &tmK, g * GF, ch.row[c], bar);
          tma_load_2d(base + OFF_K + PANEL, &tmK, g * GF + 64, ch.row[c], bar);
          tma_load_2d(base + OFF_V, &tmV, g * GF, ch.row[c], bar);
          tma_load_2d(base + OFF_V + PANEL, &tmV, g * GF + 64, ch.row[c], bar);
        }
        mbar_wait(bar, c & 1);
        tc_fence_after();
#pragma unroll
        for (int kb_ = 0; kb_ < 2; ++kb_)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(tmem + T_S, make_desc(base + OFF_AQ + kb_ * PANEL + k * 32, 16, 1024),
                      make_desc(base + OFF_K + kb_ * PANEL + k * 32, 16, 1024), idesc_kk, (kb_ | k) ? 1u : 0u);
            umma_bf16(tmem + T_DP, make_desc(base + OFF_ADO + kb_ * PANEL + k * 32, 16, 1024),
                      make_desc(base + OFF_V + kb_ * PANEL + k * 32, 16, 1024), idesc_kk, (kb_ | k) ? 1u : 0u);
          }
        umma_commit(smem_addr(&s_full));
        mbar_wait(smem_addr(&ds_full), c & 1);         // P and dS of this chunk are in shared memory, S / dP consumed
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {                   // K dim = 128 keys (dQ) / 128 (hl,t) rows (dK, dV), 16 per step
          // dQ' += dS K       A: dS K-major (k-block = k>>2, 32 B per step)      B: K tile MN-major (16 key rows per step)
          umma_bf16(tmem + T_DQ, make_desc(base + OFF_DS + (k >> 2) * PANEL + (k & 3) * 32, 16, 1024),
                    make_desc(base + OFF_K + k * 2048, PANEL, 1024), idesc_kn, (c | k) ? 1u : 0u);
          // dK = dS^T A'(Q)   A: dS MN-major (M = keys: panels of 64 keys, K = rows)   B: A'(Q) MN-major (K = rows, N = features)
          umma_bf16(tmem + T_DK, make_desc(base + OFF_DS + k * 2048, PANEL, 1024),
                    make_desc(base + OFF_AQ + k * 2048, PANEL, 1024), idesc_nn, k ? 1u : 0u);
          // dV = P^T A'(dO)
          umma_bf16(tmem + T_DV, make_desc(base + OFF_P + k * 2048, PANEL, 1024),
                    make_desc(base + OFF_ADO + k * 2048, PANEL, 1024), idesc_nn, k ? 1u : 0u);
        }
        umma_commit(smem_addr(&g_full));
      }
    }
  } else if (warp < 4) {
    const int t = lane, m = warp * 32 + lane;
    const bool live = t < a.Lq;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t any = 0;
    for (int c = 0; c < nch; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) any |= row_bits(kb, ch, c, j, t, a.causal);
    const bool rowfilled = any == 0;
    // delta = dO . O over the own head's 32 features; row statistics
    float delta = 0.f, mx = 0.f, inv = 0.f;
    if (live) {
      const __nv_bfloat16* orow = a.ctx + ((long)b * a.Lq + t) * a.ldo + g * GF + warp * DH;
      const __nv_bfloat16* grow = a.d_ctx + ((long)b * a.Lq + t) * a.ldo + g * GF + warp * DH;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float o[8], d[8];
        Act<__nv_bfloat16>::load8(orow + q * 8, o);
        Act<__nv_bfloat16>::load8(grow + q * 8, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) delta = fmaf(o[i], d[i], delta);
      }
      const float* st = a.stats + (((long)b * a.H + g * HG + warp) * a.Lq + t) * 2;
      mx = st[0]; inv = 1.f / st[1];
    }
    const float k2 = a.scale * kLog2e, m2 = mx * kLog2e;
    for (int c = 0; c < nch; ++c) {
      if (lane == 0) mbar_wait(smem_addr(&s_full), c & 1);
      __syncwarp();
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const uint32_t vb = row_bits(kb, ch, c, j, t, a.causal), eb = kb.exist[c * 4 + j];
        float pv[32], dsv[32];
        if (__any_sync(0xffffffffu, vb != 0 && !rowfilled)) {
          uint32_t rs[32], rp[32];
          tmem_ld32(trow + T_S + j * 32, rs);
          tmem_ld32(trow + T_DP + j * 32, rp);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = live && !rowfilled && ((vb >> i) & 1);
            const float p = ok ? fast_exp2(fmaf(__uint_as_float(rs[i]), k2, -m2)) * inv
                               : ((live && rowfilled && ((eb >> i) & 1)) ? inv : 0.f);
            pv[i] = p;
            dsv[i] = ok ? p * (__uint_as_float(rp[i]) - delta) * a.scale : 0.f;      // masked_fill blocks the gradient
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) { pv[i] = (live && rowfilled && ((eb >> i) & 1)) ? inv : 0.f; dsv[i] = 0.f; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v, w;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
          __nv_bfloat162* hw = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[i] = __floats2bfloat162_rn(pv[q * 8 + 2 * i], pv[q * 8 + 2 * i + 1]);
            hw[i] = __floats2bfloat162_rn(dsv[q * 8 + 2 * i], dsv[q * 8 + 2 * i + 1]);
          }
          const int key = j * 32 + q * 8;
          const uint32_t off = (key >> 6) * PANEL + sw128_offset(m, (key & 63) >> 3);
          *reinterpret_cast<uint4*>(sm + OFF_P + off) = v;
          *reinterpret_cast<uint4*>(sm + OFF_DS + off) = w;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(smem_addr(&ds_full));
      // ---- dK, dV rows of this chunk: thread = key row (warp*32 + lane), 128 features of the group
      if (lane == 0) mbar_wait(smem_addr(&g_full), c & 1);
      __syncwarp();
      tc_fence_after();
      const int ki = warp * 32 + lane;                 // key row of this chunk
#pragma unroll 1
      for (int j = 0; j < GF / 32; ++j) {
        uint32_t rk[32], rv[32];
        tmem_ld32(trow + T_DK + j * 32, rk);
        tmem_ld32(trow + T_DV + j * 32, rv);
        if (ki < ch.n[c]) {
          __nv_bfloat16* kd = a.dk + (long)(ch.row[c] + ki) * a.lddk + g * GF + j * 32;
          __nv_bfloat16* vd = a.dv + (long)(ch.row[c] + ki) * a.lddv + g * GF + j * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float x[8], y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { x[i] = __uint_as_float(rk[q * 8 + i]); y[i] = __uint_as_float(rv[q * 8 + i]); }
            Act<__nv_bfloat16>::store8(kd + q * 8, x);
            Act<__nv_bfloat16>::store8(vd + q * 8, y);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_addr(&epi_done));
    }
    // ---- dQ: own head's 32 columns (the scale is already folded into dS)
    uint32_t r[32];
    if (nch > 0) tmem_ld32(trow + T_DQ + warp * DH, r);
    if (live) {
      __nv_bfloat16* dst = a.dq + ((long)b * a.Lq + t) * a.lddq + g * GF + warp * DH;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = nch > 0 ? __uint_as_float(r[q * 8 + i]) : 0.f;
        Act<__nv_bfloat16>::store8(dst + q * 8, o);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// max_chunks: 128-key chunks a commit needs (padded batches ceil(Lk / 128); packed batches every range starts its own
// chunk, the caller passes the maximum over its commits); Lk = mask pitch
inline bool eligible(int B, int H, int Lq, int Lk, int d_head, long ldk, long ldv, int max_chunks) {
  return d_head == DH && H == 2 * HG && Lq >= 1 && Lq <= 32 && Lk >= 1 && Lk <= MAX_CH * KC && max_chunks <= MAX_CH &&
         (ldk % 8) == 0 && (ldv % 8) == 0;
}

}  // namespace attn_tc

// called by fira_attn_fwd / fira_attn_bwd (attention.cu) for the bf16 mode; returns FIRA_OK or an error code
int fira_attn_tc_fwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                     const unsigned char* key_mask, const int* ranges, long kv_rows, int causal, void* ctx, long ldo,
                     float* stats, int B, int H, int Lq, int Lk, void* stream) {
  using namespace attn_tc;
  CUtensorMap tk, tv;
  int rc = tc::make_map_bf16(&tk, k, kv_rows, 2 * GF, ldk, 64, KC, "attn_tc_fwd");
  if (rc) return rc;
  if ((rc = tc::make_map_bf16(&tv, v, kv_rows, 2 * GF, ldv, 64, KC, "attn_tc_fwd"))) return rc;
  Args a{};
  a.q = (const __nv_bfloat16*)q; a.ldq = ldq; a.key_mask = key_mask; a.ranges = ranges; a.causal = causal; a.B = B; a.H = H; a.Lq = Lq;
  a.Lk = Lk; a.scale = 1.f / sqrtf((float)DH); a.ctx = (__nv_bfloat16*)ctx; a.ldo = ldo; a.stats = stats;
  const size_t smem = (1 + MAX_CH + 2) * (size_t)TILE + 1024;
  cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { fira_set_error(FIRA_ERR_CUDA, "attn_tc_fwd attr: %s", cudaGetErrorString(e)); return FIRA_ERR_CUDA; }
  launch_k(attn_tc_fwd_kernel, dim3(B * 2), dim3(THREADS), smem, (cudaStream_t)stream, tk, tv, a);
  FIRA_CHECK_LAUNCH("fira_attn_fwd (tcgen05)");
  return FIRA_OK;
}

int fira_attn_tc_bwd(const void* q, long ldq, const void* k, long ldk, const void* v, long ldv,
                     const unsigned char* key_mask, const int* ranges, long kv_rows, int causal, const void* ctx,
                     const void* d_ctx, long ldo, const float* stats, void* dq, long lddq, void* dk, long lddk, void* dv,
                     long lddv, int B, int H, int Lq, int Lk, void* stream) {
  using namespace attn_tc;
  CUtensorMap tk, tv;
  int rc = tc::make_map_bf16(&tk, k, kv_rows, 2 * GF, ldk, 64, KC, "attn_tc_bwd");
  if (rc) return rc;
  if ((rc = tc::make_map_bf16(&tv, v, kv_rows, 2 * GF, ldv, 64, KC, "attn_tc_bwd"))) return rc;
  Args a{};
  a.q = (const __nv_bfloat16*)q; a.ldq = ldq; a.key_mask = key_mask; a.ranges = ranges; a.causal = causal; a.B = B; a.H = H; a.Lq = Lq;
  a.Lk = Lk; a.scale = 1.f / sqrtf((float)DH); a.ctx = (__nv_bfloat16*)const_cast<void*>(ctx); a.ldo = ldo;
  a.stats = const_cast<float*>(stats); a.d_ctx = (const __nv_bfloat16*)d_ctx;
  a.dq = (__nv_bfloat16*)dq; a.lddq = lddq; a.dk = (__nv_bfloat16*)dk; a.lddk = lddk; a.dv = (__nv_bfloat16*)dv; a.lddv = lddv;
  const size_t smem = 6 * (size_t)TILE + 1024;
  cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { fira_set_error(FIRA_ERR_CUDA, "attn_tc_bwd attr: %s", cudaGetErrorString(e)); return FIRA_ERR_CUDA; }
  launch_k(attn_tc_bwd_kernel, dim3(B * 2), dim3(THREADS), smem, (cudaStream_t)stream, tk, tv, a);
  FIRA_CHECK_LAUNCH("fira_attn_bwd (tcgen05)");
  return FIRA_OK;
}

bool fira_attn_tc_eligible(int B, int H, int Lq, int Lk, int d_head, long ldk, long ldv, int max_chunks) {
  return attn_tc::eligible(B, H, Lq, Lk, d_head, ldk, ldv, max_chunks);
}
