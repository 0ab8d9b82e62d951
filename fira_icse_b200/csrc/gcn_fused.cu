// Fused GCN layer for the bf16 throughput path (gnn_transformer.py:74-86), ONE kernel per layer and direction:
//
//   forward  (MODE 0):  Z = (A H) Wc^T + rowsum(A) (x) c1 + b2 ;  out = LN(dropout(Z) + H)          Wc = W2 W1, c1 = W2 b1
//   backward (MODE 1):  AdZ = A^T dZ (side output for dWc = AdZ^T H) ;  dH = AdZ Wc + d_resid
//
// gather -> transform -> epilogue without a round trip through HBM between them:
//   * 8 gather warps build the aggregated 128-row tile (A H)[rows, 0:256] straight in shared memory, in the
//     K-major SWIZZLE_128B layout tcgen05.mma reads (the tile IS the A operand): a quarter warp per destination row,
//     16-byte loads, 8 x 128 B of neighbour rows in flight per lane group, fp32 accumulation in CSR order, the tile's
//     rowptr / (col, val) metadata prefetched into shared memory first (one dependent latency per row, not three);
//   * the 256 x 256 weight (128 KB bf16) is loaded ONCE per CTA by TMA and stays resident (persistent CTAs, one per SM);
//   * one elected thread issues 16 tcgen05.mma (M128 N256 K16) per tile into one of TWO 256-column TMEM accumulators,
//     so the epilogue of tile t overlaps the gather of tile t+1;
//   * 4 epilogue warps read the accumulator (tcgen05.ld: lane = row), add bias + rowsum*c1, write Z, apply dropout +
//     residual + LayerNorm (row statistics are thread-local: a thread owns a whole row) and write the normalised rows;
//     all global traffic of the epilogue goes through a small per-warp staging block so loads/stores are 64-B row
//     segments instead of one row per lane.
// Rows are addressed through a CSR in BUFFER order (rowptr[r], col = buffer row; fira_csr_to_rows builds it from the
// (graph, node)-ordered CSR), so the kernel does not care whether the node buffer is padded segment-major or packed.
#include "tc_common.cuh"
#include "fira_b200.h"

namespace {

using namespace tc;

constexpr int D = 256;
constexpr int TM = 128;                    // rows per tile = UMMA M
constexpr int N_EPI = 8, N_GATHER = 8;     // epilogue warp w: TMEM lane quarter w & 3, column half w >> 2
constexpr int WARP_MMA = N_EPI;            // warps 0-7 epilogue, 8 = TMA + MMA, 9-16 gather
constexpr int THREADS = (N_EPI + 1 + N_GATHER) * 32;
constexpr int EC = 1024;                   // edges of a tile staged in shared memory (larger tiles read col/val from global)
constexpr int STG_PITCH = 80;              // bytes per staged row of 32 bf16 (64 B) + 16 B pad: conflict-free 16-B accesses
constexpr uint32_t B_BYTES = D * D * 2;    // 131072: 4 k-blocks x [256 n-rows x 128 B]
constexpr uint32_t A_BYTES = TM * D * 2;   // 65536:  4 k-blocks x [128 rows x 128 B]
constexpr uint32_t STG_BYTES = N_EPI * 32 * STG_PITCH;          // per epilogue warp: one [32 x 32] bf16 block (in, then out)
constexpr uint32_t OFF_A = B_BYTES, OFF_STG = OFF_A + A_BYTES, OFF_ROWPTR = OFF_STG + STG_BYTES,
                   OFF_COL = OFF_ROWPTR + 544, OFF_VAL = OFF_COL + EC * 4, OFF_RS = OFF_VAL + EC * 4,
                   OFF_ST = OFF_RS + 4 * TM * 4, SMEM_BYTES = OFF_ST + 2 * TM * 2 * 4;
constexpr int GATHER_BAR = 1;              // named barrier of the gather warps
constexpr int EPI_BAR = 2;                 // named barrier of the epilogue warps (row statistics exchange)

struct Params {
  const int* rowptr; const int* col; const float* val;      // buffer-order CSR
  const __nv_bfloat16* x;                                    // H (fwd) or dZ (bwd), [R, 256]
  long R;
  // MODE 0
  const float* bias; const float* c1; const float* gamma; const float* beta;
  __nv_bfloat16* z; __nv_bfloat16* outA; __nv_bfloat16* outB; long split;
  float* mean; float* rstd;
  float p_drop; uint64_t seed; const uint64_t* seed_ctr; uint32_t stream_id;
  // MODE 1
  const __nv_bfloat16* addend; __nv_bfloat16* agg_out; __nv_bfloat16* y;
};

__device__ __forceinline__ void tile_range(long R, int cta, int ncta, int t, long& r0, int& rows, int& ntiles) {
  const long rpc = (R + ncta - 1) / ncta;
  const long lo = (long)cta * rpc;
  const long hi = lo + rpc < R ? lo + rpc : R;
  const long n = hi > lo ? hi - lo : 0;
  ntiles = (int)((n + TM - 1) / TM);
  if (ntiles == 0) { r0 = 0; rows = 0; return; }
  long chunk = (n + ntiles - 1) / ntiles;
  chunk = (chunk + 7) & ~7L;
  if (chunk > TM) chunk = TM;
  r0 = lo + (long)t * chunk;
  const long e = r0 + chunk < hi ? r0 + chunk : hi;
  rows = (int)(e > r0 ? e - r0 : 0);
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return r;
}
__device__ __forceinline__ void unpack8(const uint4& raw, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
// 32 lanes x 32 consecutive fp32 columns back into TMEM (the layout tmem_ld32 reads)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) gcn_fused_kernel(const __grid_constant__ CUtensorMap tmW, Params p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long b_full, a_full, a_empty, tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  unsigned char* sm = smem_raw + (base - smem_addr(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  long r0_; int rows_, ntiles;
  tile_range(p.R, cta, ncta, 0, r0_, rows_, ntiles);

  if (threadIdx.x == 0) {
    mbar_init(smem_addr(&b_full), 1);
    mbar_init(smem_addr(&a_full), N_GATHER * 32);
    mbar_init(smem_addr(&a_empty), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_addr(&tmem_full[i]), 1); mbar_init(smem_addr(&tmem_empty[i]), N_EPI * 32); }
    mbar_init_fence();
    tma_prefetch_desc(&tmW);
  }
  if (warp == WARP_MMA) tmem_alloc(smem_addr(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait(); pdl_trigger();       // PDL: the prologue above overlapped the previous kernel's tail (common.cuh)

  if (warp == WARP_MMA) {
    if (lane == 0 && ntiles > 0) {
      // ---- weight: resident B operand, 4 k-blocks of [256 n x 64 k]
      mbar_expect_tx(smem_addr(&b_full), B_BYTES);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(base + kb * 32768, &tmW, kb * 64, 0, smem_addr(&b_full));
      mbar_wait(smem_addr(&b_full), 0);
      constexpr uint32_t idesc = make_idesc_bf16(TM, D, false, false);
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t >= 2) mbar_wait(smem_addr(&tmem_empty[buf]), ((t >> 1) - 1) & 1);      // epilogue of tile t-2 drained it
        mbar_wait(smem_addr(&a_full), t & 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_desc(base + OFF_A + kb * 16384 + k * 32, 16, 1024);
            const uint64_t db = make_desc(base + kb * 32768 + k * 32, 16, 1024);
            umma_bf16(tmem_base + buf * D, da, db, idesc, (kb | k) ? 1u : 0u);
          }
        umma_commit(smem_addr(&a_empty));            // A tile may be overwritten
        umma_commit(smem_addr(&tmem_full[buf]));     // accumulator ready
      }
    }
  } else if (warp > WARP_MMA) {
    // ======================================================= gather warps: A tile = (A_hat X)[tile rows, :]
    const int gw = warp - WARP_MMA - 1;              // 0..7
    const int gtid = gw * 32 + lane;                 // 0..255
    const int q = lane >> 3, ql = lane & 7;          // quarter-warp (row slot) / lane within it
    int* s_rowptr = reinterpret_cast<int*>(sm + OFF_ROWPTR);
    int* s_col = reinterpret_cast<int*>(sm + OFF_COL);
    float* s_val = reinterpret_cast<float*>(sm + OFF_VAL);
    float* s_rs = reinterpret_cast<float*>(sm + OFF_RS);
    for (int t = 0; t < ntiles; ++t) {
      long r0; int rows, nt_;
      tile_range(p.R, cta, ncta, t, r0, rows, nt_);
      // ---- tile metadata -> shared memory (every gather warp is done with the previous tile's metadata first)
      if (t >= 1) asm volatile("bar.sync %0, %1;" ::"n"(GATHER_BAR), "n"(N_GATHER * 32) : "memory");
      for (int i = gtid; i <= rows; i += N_GATHER * 32) s_rowptr[i] = p.rowptr[r0 + i];
      asm volatile("bar.sync %0, %1;" ::"n"(GATHER_BAR), "n"(N_GATHER * 32) : "memory");
      const int e_lo = s_rowptr[0], e_hi = s_rowptr[rows];
      const int nE = e_hi - e_lo;
      const bool staged = nE <= EC;
      if (staged)
        for (int i = gtid; i < nE; i += N_GATHER * 32) { s_col[i] = p.col[e_lo + i]; s_val[i] = p.val[e_lo + i]; }
      if (t >= 1 && gtid == 0) mbar_wait(smem_addr(&a_empty), (t - 1) & 1);   // MMAs of tile t-1 have read the A tile
      asm volatile("bar.sync %0, %1;" ::"n"(GATHER_BAR), "n"(N_GATHER * 32) : "memory");
      const int* cp = staged ? s_col : p.col + e_lo;
      const float* vp = staged ? s_val : p.val + e_lo;
      // each warp: rows gw*16 + it*4 + q
#pragma unroll 1
      for (int it = 0; it < 4; ++it) {
        const int r = gw * 16 + it * 4 + q;
        if (r < rows) {
          const int e0 = s_rowptr[r] - e_lo, e1 = s_rowptr[r + 1] - e_lo;
          float acc[4][8];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
          float rsum = 0.f;
          int e = e0;
          for (; e + 1 < e1; e += 2) {
            const int c0 = cp[e], c1 = cp[e + 1];
            const float w0 = vp[e], w1 = vp[e + 1];
            const uint4* s0 = reinterpret_cast<const uint4*>(p.x + (long)c0 * D + ql * 8);
            const uint4* s1 = reinterpret_cast<const uint4*>(p.x + (long)c1 * D + ql * 8);
            uint4 a[4], b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = __ldg(s0 + j * 8);       // features j*64 + ql*8 .. +7
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = __ldg(s1 + j * 8);
            rsum += w0 + w1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
              unpack8(a[j], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(w0, f[i], acc[j][i]);
              unpack8(b[j], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(w1, f[i], acc[j][i]);
            }
          }
          if (e < e1) {
            const int c0 = cp[e];
            const float w0 = vp[e];
            const uint4* s0 = reinterpret_cast<const uint4*>(p.x + (long)c0 * D + ql * 8);
            uint4 a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = __ldg(s0 + j * 8);
            rsum += w0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
              unpack8(a[j], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(w0, f[i], acc[j][i]);
            }
          }
          // bf16 row -> swizzled K-major A tile: k-block j, 16-byte chunk ql of row r
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 v = pack8(acc[j]);
            *reinterpret_cast<uint4*>(sm + OFF_A + j * 16384 + sw128_offset(r, ql)) = v;
            if (MODE == 1) *reinterpret_cast<uint4*>(p.agg_out + (r0 + r) * D + j * 64 + ql * 8) = v;
          }
          if (MODE == 0 && ql == 0) s_rs[(t & 3) * TM + r] = rsum;
        }
      }
      fence_proxy_async();                           // my generic-proxy writes -> visible to the UMMA (async proxy) reads
      mbar_arrive(smem_addr(&a_full));
    }
  } else {
    // ======================================================= epilogue warps: TMEM lanes 32*(warp & 3), columns 128*(warp >> 2)
    // Per warp and 32-column chunk: the residual / addend block [32 rows x 32 cols] is fetched one chunk AHEAD into
    // registers (coalesced 64-B row segments), handed over through a staging block (lane = row afterwards); results go
    // back through a second staging block so that global stores are 64-B row segments too.  MODE 0 keeps
    // y = dropout(z) + h in TMEM (tcgen05.st) between the statistics pass and the normalisation pass.
    const int quarter = warp & 3, half = warp >> 2;
    unsigned char* stg = sm + OFF_STG + warp * 32 * STG_PITCH;
    const float* s_rs = reinterpret_cast<const float*>(sm + OFF_RS);
    float* s_st = reinterpret_cast<float*>(sm + OFF_ST);                 // [2 halves][128 rows][sum, sumsq]
    uint64_t seed = p.seed;
    if (MODE == 0 && p.seed_ctr) seed += *p.seed_ctr;
    const float keep_scale = (MODE == 0 && p.p_drop > 0.f) ? 1.f / (1.f - p.p_drop) : 1.f;
    const int lrow = lane >> 2, lch = lane & 3;      // cooperative 64-B row segments: 8 rows x 4 chunks per instruction
    const __nv_bfloat16* resid = MODE == 0 ? p.x : p.addend;
    for (int t = 0; t < ntiles; ++t) {
      long r0; int rows, nt_;
      tile_range(p.R, cta, ncta, t, r0, rows, nt_);
      const int buf = t & 1;
      const int wrow0 = quarter * 32;                // first tile row of this warp
      const int my = wrow0 + lane;                   // this thread's tile row
      const bool live = my < rows;
      const long grow = r0 + my;
      const uint32_t tacc = tmem_base + buf * D + ((uint32_t)wrow0 << 16) + half * 128;
      const int col0 = half * 128;
      // residual / addend: every thread reads ITS row directly (16-B loads, two 32-column chunks ahead in registers);
      // the first two chunks are requested before the accumulator is ready, so their latency hides behind the MMAs
      uint4 hq[2][4];
      auto fetch = [&](int k, uint4* dst) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dst[j] = make_uint4(0, 0, 0, 0);
          if (live && resid != nullptr) dst[j] = __ldg(reinterpret_cast<const uint4*>(resid + grow * D + col0 + k * 32 + j * 8));
        }
      };
      fetch(0, hq[0]);
      fetch(1, hq[1]);
      if (lane == 0) mbar_wait(smem_addr(&tmem_full[buf]), (t >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      float rs = 0.f, sum = 0.f, sq = 0.f;
      if (MODE == 0 && live) rs = s_rs[(t & 3) * TM + my];
      auto chunk = [&](int k, const uint4* hp) {
        const int c = col0 + k * 32;
        uint32_t acc[32];
        tmem_ld32(tacc + k * 32, acc);
        float y[32];
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + c + j));
            const float4 cv = __ldg(reinterpret_cast<const float4*>(p.c1 + c + j));
            y[j + 0] = fmaf(rs, cv.x, __uint_as_float(acc[j + 0]) + bv.x);
            y[j + 1] = fmaf(rs, cv.y, __uint_as_float(acc[j + 1]) + bv.y);
            y[j + 2] = fmaf(rs, cv.z, __uint_as_float(acc[j + 2]) + bv.z);
            y[j + 3] = fmaf(rs, cv.w, __uint_as_float(acc[j + 3]) + bv.w);
          }
          // Z is stored as bf16 and the LayerNorm backward recomputes from the stored value: normalise the same value
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 zp = pack8(y + j * 8);
            *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + j * 16) = zp;
            unpack8(zp, y + j * 8);
          }
          if (p.p_drop > 0.f) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t m = dropout_keep8(seed, p.stream_id, (uint64_t)grow * 32 + (c >> 3) + j, p.p_drop);
#pragma unroll
              for (int i = 0; i < 8; ++i) y[j * 8 + i] = ((m >> i) & 1) ? y[j * 8 + i] * keep_scale : 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float h[8];
            unpack8(hp[j], h);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float v = y[j * 8 + i] + h[i]; y[j * 8 + i] = v; sum += v; sq = fmaf(v, v, sq); }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(y[j]);
          tmem_st32(tacc + k * 32, acc);             // y stays in TMEM for the normalisation pass
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float h[8];
            unpack8(hp[j], h);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[j * 8 + i] = __uint_as_float(acc[j * 8 + i]) + h[i];
            *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + j * 16) = pack8(y + j * 8);
          }
        }
        __syncwarp();
        // staged [32 x 32] bf16 block (Z, or dH) -> global, 64-B row segments
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + lrow;
          if (wrow0 + rr < rows) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * STG_PITCH + lch * 16);
            __nv_bfloat16* dst = MODE == 0 ? p.z : p.y;
            *reinterpret_cast<uint4*>(dst + (r0 + wrow0 + rr) * D + c + lch * 8) = v;
          }
        }
        __syncwarp();
      };
#pragma unroll 1
      for (int k = 0; k < 4; k += 2) {
        chunk(k, hq[0]);
        if (k + 2 < 4) fetch(k + 2, hq[0]);
        chunk(k + 1, hq[1]);
        if (k + 3 < 4) fetch(k + 3, hq[1]);
      }
      if (MODE == 0) {
        tmem_st_wait();
        // row statistics over all 256 columns: the two warps of a lane quarter exchange their halves
        s_st[(half * TM + my) * 2] = sum;
        s_st[(half * TM + my) * 2 + 1] = sq;
        asm volatile("bar.sync %0, %1;" ::"n"(EPI_BAR), "n"(N_EPI * 32) : "memory");
        const float tsum = sum + s_st[((half ^ 1) * TM + my) * 2];
        const float tsq = sq + s_st[((half ^ 1) * TM + my) * 2 + 1];
        const float mean = tsum * (1.f / D);
        const float var = fmaxf(tsq * (1.f / D) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + kLnEps);
        if (live && half == 0 && p.mean) { p.mean[grow] = mean; p.rstd[grow] = rstd; }
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int c = col0 + k * 32;
          uint32_t acc[32];
          tmem_ld32(tacc + k * 32, acc);
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(p.gamma + c + j));
            const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + c + j));
            o[j + 0] = fmaf((__uint_as_float(acc[j + 0]) - mean) * rstd, gv.x, bt.x);
            o[j + 1] = fmaf((__uint_as_float(acc[j + 1]) - mean) * rstd, gv.y, bt.y);
            o[j + 2] = fmaf((__uint_as_float(acc[j + 2]) - mean) * rstd, gv.z, bt.z);
            o[j + 3] = fmaf((__uint_as_float(acc[j + 3]) - mean) * rstd, gv.w, bt.w);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + j * 16) = pack8(o + j * 8);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + lrow;
            if (wrow0 + rr < rows) {
              const long gr = r0 + wrow0 + rr;
              const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * STG_PITCH + lch * 16);
              __nv_bfloat16* dst = gr < p.split ? p.outA : p.outB;
              *reinterpret_cast<uint4*>(dst + gr * D + c + lch * 8) = v;
            }
          }
          __syncwarp();
        }
        // the partner warp must have read s_st before the next tile overwrites it
        asm volatile("bar.sync %0, %1;" ::"n"(EPI_BAR), "n"(N_EPI * 32) : "memory");
      }
      tc_fence_before();
      mbar_arrive(smem_addr(&tmem_empty[buf]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- (graph, node)-ordered CSR -> buffer-order CSR
struct Segs { int B, n0, n1, n2; };
__device__ __forceinline__ long seg_row(const Segs& s, int b, int j) {
  if (j < s.n0) return (long)b * s.n0 + j;
  if (j < s.n0 + s.n1) return (long)s.B * s.n0 + (long)b * s.n1 + (j - s.n0);
  return (long)s.B * (s.n0 + s.n1) + (long)b * s.n2 + (j - s.n0 - s.n1);
}
__device__ __forceinline__ void seg_unrow(const Segs& s, long r, int& b, int& i) {
  const long e0 = (long)s.B * s.n0, e1 = e0 + (long)s.B * s.n1;
  if (r < e0) { b = (int)(r / s.n0); i = (int)(r % s.n0); }
  else if (r < e1) { long q = r - e0; b = (int)(q / s.n1); i = s.n0 + (int)(q % s.n1); }
  else { long q = r - e1; b = (int)(q / s.n2); i = s.n0 + s.n1 + (int)(q % s.n2); }
}

__global__ void rows_count_kernel(const int* __restrict__ rowptr, Segs s, int N, int* __restrict__ counts) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  const long R = (long)s.B * N;
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (long)gridDim.x * blockDim.x) {
    int b, i; seg_unrow(s, r, b, i);
    const long g = (long)b * N + i;
    counts[r] = rowptr[g + 1] - rowptr[g];
  }
}

// one 1024-thread CTA: exclusive scan of counts[0..n) -> out[0..n]
__global__ void rows_scan_kernel(const int* __restrict__ counts, int* __restrict__ out, long n) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long b0 = 0; b0 < n; b0 += 1024) {
    const long i = b0 + threadIdx.x;
    const int v = i < n ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + x - v;
    if (i < n) out[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

__global__ void rows_fill_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ val,
                                 Segs s, int N, const int* __restrict__ rowptr_g, int* __restrict__ col_g,
                                 float* __restrict__ val_g) {
  pdl_wait(); pdl_trigger();       // PDL (common.cuh)
  const long R = (long)s.B * N;
  const int lane = threadIdx.x & 31;
  for (long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < R; r += (long)gridDim.x * (blockDim.x >> 5)) {
    int b, i; seg_unrow(s, r, b, i);
    const long g = (long)b * N + i;
    const int e0 = rowptr[g], n = rowptr[g + 1] - e0, o0 = rowptr_g[r];
    for (int k = lane; k < n; k += 32) { col_g[o0 + k] = (int)seg_row(s, b, col[e0 + k]); val_g[o0 + k] = val[e0 + k]; }
  }
}

int launch_fused(int mode, const CUtensorMap& tm, const Params& p, cudaStream_t st) {
  static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "fused GCN kernel: shared-memory budget");
  const size_t smem = SMEM_BYTES + 1024;
  cudaError_t e = mode == 0
      ? cudaFuncSetAttribute(gcn_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
      : cudaFuncSetAttribute(gcn_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { fira_set_error(FIRA_ERR_CUDA, "gcn_layer attr: %s", cudaGetErrorString(e)); return FIRA_ERR_CUDA; }
  long want = (p.R + 31) / 32;
  const int grid = (int)(want < 148 ? (want < 1 ? 1 : want) : 148);
  if (mode == 0) launch_k(gcn_fused_kernel<0>, dim3(grid), dim3(THREADS), smem, st, tm, p);
  else launch_k(gcn_fused_kernel<1>, dim3(grid), dim3(THREADS), smem, st, tm, p);
  return FIRA_OK;
}

}  // namespace

extern "C" {

int fira_csr_to_rows(const int* rowptr, const int* col, const float* val, int B, int n_code, int n_sub, int n_ast,
                     int* counts, int* rowptr_rows, int* col_rows, float* val_rows, void* stream) {
  FIRA_CHECK_ARG(rowptr && col && val && counts && rowptr_rows && col_rows && val_rows, FIRA_ERR_ARG, "csr_to_rows: null");
  FIRA_CHECK_ARG(B > 0 && n_code > 0 && n_sub >= 0 && n_ast >= 0, FIRA_ERR_SHAPE, "csr_to_rows: segments");
  Segs s{B, n_code, n_sub, n_ast};
  const int N = n_code + n_sub + n_ast;
  const long R = (long)B * N;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (int)((R + 255) / 256 < 148 * 4 ? (R + 255) / 256 : 148 * 4);
  launch_k(rows_count_kernel, dim3(grid), dim3(256), 0, st, rowptr, s, N, counts);
  launch_k(rows_scan_kernel, dim3(1), dim3(1024), 0, st, counts, rowptr_rows, R);
  grid = (int)((R + 7) / 8 < 148 * 8 ? (R + 7) / 8 : 148 * 8);
  launch_k(rows_fill_kernel, dim3(grid), dim3(256), 0, st, rowptr, col, val, s, N, rowptr_rows, col_rows, val_rows);
  FIRA_CHECK_LAUNCH("fira_csr_to_rows");
  return FIRA_OK;
}

int fira_gcn_layer_fwd(const int* rowptr_rows, const int* col_rows, const float* val_rows, const void* h,
                       const void* w_merged, const float* bias, const float* c1, const float* gamma, const float* beta,
                       void* z, void* outA, void* outB, long split, float* mean, float* rstd, long rows, int dim,
                       float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t stream_id, void* stream) {
  FIRA_CHECK_ARG(dim == D, FIRA_ERR_SHAPE, "gcn_layer_fwd: dim %d != 256", dim);
  FIRA_CHECK_ARG(rows > 0, FIRA_ERR_SHAPE, "gcn_layer_fwd: rows %ld", rows);
  FIRA_CHECK_ARG(rowptr_rows && col_rows && val_rows && h && w_merged && bias && c1 && gamma && beta && z && outA && outB,
                 FIRA_ERR_ARG, "gcn_layer_fwd: null argument");
  FIRA_CHECK_ARG(fira_aligned16(h) && fira_aligned16(w_merged) && fira_aligned16(z) && fira_aligned16(outA) &&
                     fira_aligned16(outB) && fira_aligned16(bias) && fira_aligned16(c1) && fira_aligned16(gamma) &&
                     fira_aligned16(beta), FIRA_ERR_ALIGN, "gcn_layer_fwd: 16-B alignment");
  FIRA_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, FIRA_ERR_ARG, "gcn_layer_fwd: p_drop %f", p_drop);
  CUtensorMap tm;
  int rc = make_map_bf16(&tm, w_merged, D, D, D, 64, D, "gcn_layer_fwd");
  if (rc) return rc;
  Params p{};
  p.rowptr = rowptr_rows; p.col = col_rows; p.val = val_rows; p.x = (const __nv_bfloat16*)h; p.R = rows;
  p.bias = bias; p.c1 = c1; p.gamma = gamma; p.beta = beta;
  p.z = (__nv_bfloat16*)z; p.outA = (__nv_bfloat16*)outA; p.outB = (__nv_bfloat16*)outB; p.split = split;
  p.mean = mean; p.rstd = rstd; p.p_drop = p_drop; p.seed = seed; p.seed_ctr = seed_ctr; p.stream_id = stream_id;
  if ((rc = launch_fused(0, tm, p, (cudaStream_t)stream))) return rc;
  FIRA_CHECK_LAUNCH("fira_gcn_layer_fwd");
  return FIRA_OK;
}

int fira_gcn_layer_bwd(const int* rowptr_rows_t, const int* col_rows_t, const float* val_rows_t, const void* d_z,
                       const void* w_merged_t, const void* d_resid, void* agg_dz, void* d_h, long rows, int dim,
                       void* stream) {
  FIRA_CHECK_ARG(dim == D, FIRA_ERR_SHAPE, "gcn_layer_bwd: dim %d != 256", dim);
  FIRA_CHECK_ARG(rows > 0, FIRA_ERR_SHAPE, "gcn_layer_bwd: rows %ld", rows);
  FIRA_CHECK_ARG(rowptr_rows_t && col_rows_t && val_rows_t && d_z && w_merged_t && agg_dz && d_h, FIRA_ERR_ARG,
                 "gcn_layer_bwd: null argument");
  FIRA_CHECK_ARG(fira_aligned16(d_z) && fira_aligned16(w_merged_t) && fira_aligned16(d_resid) && fira_aligned16(agg_dz) &&
                     fira_aligned16(d_h), FIRA_ERR_ALIGN, "gcn_layer_bwd: 16-B alignment");
  CUtensorMap tm;
  int rc = make_map_bf16(&tm, w_merged_t, D, D, D, 64, D, "gcn_layer_bwd");
  if (rc) return rc;
  Params p{};
  p.rowptr = rowptr_rows_t; p.col = col_rows_t; p.val = val_rows_t; p.x = (const __nv_bfloat16*)d_z; p.R = rows;
  p.addend = (const __nv_bfloat16*)d_resid; p.agg_out = (__nv_bfloat16*)agg_dz; p.y = (__nv_bfloat16*)d_h;
  if ((rc = launch_fused(1, tm, p, (cudaStream_t)stream))) return rc;
  FIRA_CHECK_LAUNCH("fira_gcn_layer_bwd");
  return FIRA_OK;
}

}  // extern "C"
