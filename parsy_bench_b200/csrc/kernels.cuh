// Device kernels of the B200 executor (sm_100a).  FP64 throughout.
//
//   k_gemm_tiles     DMMA (mma.sync.m8n8k4.f64) tile kernel: SYRK/GEMM update with fused scatter-subtract
//                    epilogue (K3+K4+K5 of SURVEY.md §2c) and TRSM through the inverse diagonal block (K7).
//   k_update_small   warp-cooperative FMA update for narrow pairs (K <= 32, ndrow1 <= 32).
//   k_factor_small   one warp per narrow supernode: POTRF + TRSM fused (K6+K7).
//   k_potrf_block    one CTA per block column: POTRF of the <=128-wide diagonal block in shared memory
//                    + its inverse (feeds the DMMA TRSM and the diagonal solves of the sweeps).
//   k_fwd_* / k_bwd_*  supernodal forward / backward sweeps (K8-K10).
//   k_build_rel / k_build_apos / k_assemble   structure-time helpers and the A -> L scatter (K1).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "plan.h"

namespace parsy {

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 8 : 0;   // src-size 0 => zero-fill, nothing is read
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
// same with a 32-bit shared-window address computed by the caller (hoisted out of the K loop)
__device__ __forceinline__ void cp_async8_s(unsigned saddr, const void* gmem, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(saddr), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(8x8) += A(8x4, row) * B(4x8, col); lane holds A[lane>>2][lane&3], B[lane&3][lane>>2],
// C[lane>>2][(lane&3)*2 + {0,1}].  SASS: DMMA.8x8x4 (the only native FP64 tensor shape on sm_100a).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------
// DMMA tile kernel
// ------------------------------------------------------------------------------------------------
template <int TM_, int TN_, int WM_, int WN_, int KC_, int STAGES_>
struct GemmCfg {
  static constexpr int TM = TM_, TN = TN_, WM = WM_, WN = WN_, KC = KC_, STAGES = STAGES_;
  static constexpr int WARPS_M = TM / WM, WARPS_N = TN / WN, THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int LDA = TM + 4, LDB = TN + 4;   // (LD mod 16) == 4 -> conflict-free fragment loads
  static constexpr int STAGE_DOUBLES = KC * (LDA + LDB);
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_DOUBLES * 8 + (TM + TN) * 4;
  static constexpr int FM = WM / 8, FN = WN / 8;
  static constexpr int MIN_CTAS = (THREADS <= 256 && SMEM <= 100 * 1024) ? 2 : 1;
};
// 128x64 tiles, 8 warps, two CTAs per SM: one CTA's prologue/epilogue overlaps the other's main loop and the SM
// still holds 16 warps (4 per SMSP) to keep the DMMA pipe fed across LDS/barrier stalls
using Cfg128 = GemmCfg<128, 64, 32, 32, 16, 3>;
using Cfg64 = GemmCfg<64, 64, 32, 32, 8, 4>;
// latency configuration for launches that cannot fill the GPU with larger tiles (the panel -> next block column
// updates on the critical path of a separator): 32x32 tiles, 4 warps of 16x16; a tile with K = 128 is ~1 us of DMMA
// issue on its SM, so the launch is spread over up to 16x more SMs than with 128x64 tiles
using Cfg32 = GemmCfg<32, 32, 16, 16, 16, 3>;
// TRSM: TN = 128 keeps it in place (a CTA owns whole rows); the row-tile height is picked per step so that the
// latency-critical panel solve covers the GPU: 64, 32 or 16 rows
using CfgTrsm = GemmCfg<64, 128, 32, 32, 16, 3>;
using CfgTrsm32 = GemmCfg<32, 128, 16, 32, 16, 3>;
using CfgTrsm16 = GemmCfg<16, 128, 16, 32, 16, 3>;

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MIN_CTAS) k_gemm_tiles(const GemmTask* __restrict__ tasks, int ntasks,
                                                            double* __restrict__ lv, const double* __restrict__ linv,
                                                            const int* __restrict__ rel, int tile_base) {
  extern __shared__ __align__(16) double smem[];
  int* srel = reinterpret_cast<int*>(smem + C::STAGES * C::STAGE_DOUBLES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bid = blockIdx.x + tile_base;   // long tile lists are launched in slices (launch_tiles in solver.cu)
  // locate the task that owns this tile (tile0 is an exclusive prefix over the launch)
  int lo = 0, hi = ntasks - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tasks[mid].tile0 <= bid) lo = mid; else hi = mid - 1;
  }
  const GemmTask T = tasks[lo];
  int t = bid - T.tile0;
  const int MT = (T.M + C::TM - 1) / C::TM;
  int mi, ni;
  if (T.flags & GF_LOWER) {        // lower trapezoid of tiles: column tile ni holds the row tiles that reach its columns
    ni = 0;
    int first = 0, cnt = MT;
    while (t >= cnt) { t -= cnt; ++ni; first = (ni * C::TN) / C::TM; cnt = MT - first; }
    mi = first + t;
  } else {
    mi = t % MT; ni = t / MT;
  }
  const int m0 = mi * C::TM, n0 = ni * C::TN;
  const int mrows = min(C::TM, T.M - m0), nrows = min(C::TN, T.N - n0);
  const double* __restrict__ A = lv + T.a_off + m0;
  const double* __restrict__ B = ((T.flags & GF_B_LINV) ? linv : lv) + T.b_off + n0;
  const int K = T.K, lda = T.lda, ldb = T.ldb;

  for (int i = tid; i < C::TM + C::TN; i += C::THREADS) {
    const bool isrow = i < C::TM;
    const int loc = isrow ? i : i - C::TM;
    const int idx = isrow ? m0 + loc : n0 + loc;
    const bool valid = isrow ? (loc < mrows) : (loc < nrows);
    int v = idx;
    if (T.rel_off >= 0) v = valid ? rel[T.rel_off + idx] : 0;
    srel[i] = v;
  }

  // Operand staging.  Every thread copies the same tile row (ia / jb) of RA / RB consecutive k-columns per pass, so the
  // row predicate, the shared-memory address and the global pointer are set up once per tile; a K-chunk is then
  // KC/RA + KC/RB copies at fixed strides (the per-element index arithmetic of a generic loop cost ~20 instructions
  // per 8-byte copy — more issue slots than the DMMAs themselves).  Only a partial last chunk tests k < K per copy.
  constexpr int RA = C::THREADS / C::TM, NA = C::KC / RA, RB = C::THREADS / C::TN, NBP = C::KC / RB;
  static_assert(C::THREADS % C::TM == 0 && C::THREADS % C::TN == 0 && C::KC % RA == 0 && C::KC % RB == 0 && NA >= 1 && NBP >= 1,
                "tile shape must give every thread a fixed operand row");
  const int ia = tid % C::TM, ka = tid / C::TM, jb = tid % C::TN, kb = tid / C::TN;
  const int szA = ia < mrows ? 8 : 0, szB = jb < nrows ? 8 : 0;     // src-size 0: zero-fill, nothing is read
  const double* pA = A + (int64_t)ka * lda + (szA ? ia : 0);
  const double* pB = B + (int64_t)kb * ldb + (szB ? jb : 0);
  const int64_t passA = (int64_t)RA * lda, passB = (int64_t)RB * ldb;
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned sA = s0 + (unsigned)(ka * C::LDA + ia) * 8u;
  const unsigned sB = s0 + (unsigned)(C::KC * C::LDA + kb * C::LDB + jb) * 8u;
  int k_next = 0;                                                    // chunks are loaded in increasing order
  auto load_chunk = [&](int /*kc*/, int stage) {
    const unsigned so = (unsigned)stage * (unsigned)(C::STAGE_DOUBLES * 8);
    if (k_next + C::KC <= K) {
#pragma unroll
      for (int u = 0; u < NA; ++u) cp_async8_s(sA + so + (unsigned)(u * RA * C::LDA * 8), pA + u * passA, szA);
#pragma unroll
      for (int u = 0; u < NBP; ++u) cp_async8_s(sB + so + (unsigned)(u * RB * C::LDB * 8), pB + u * passB, szB);
    } else {
#pragma unroll
      for (int u = 0; u < NA; ++u) {
        const bool in = k_next + ka + u * RA < K;
        cp_async8_s(sA + so + (unsigned)(u * RA * C::LDA * 8), in ? (const void*)(pA + u * passA) : (const void*)tasks, in ? szA : 0);
      }
#pragma unroll
      for (int u = 0; u < NBP; ++u) {
        const bool in = k_next + kb + u * RB < K;
        cp_async8_s(sB + so + (unsigned)(u * RB * C::LDB * 8), in ? (const void*)(pB + u * passB) : (const void*)tasks, in ? szB : 0);
      }
    }
    k_next += C::KC;
    pA += (int64_t)C::KC * lda;
    pB += (int64_t)C::KC * ldb;
  };

  const int nchunks = (K + C::KC - 1) / C::KC;
#pragma unroll
  for (int s = 0; s < C::STAGES - 1; ++s) {
    if (s < nchunks) load_chunk(s, s);
    cp_async_commit();
  }
  double acc[C::FM][C::FN][2];
#pragma unroll
  for (int im = 0; im < C::FM; ++im)
#pragma unroll
    for (int in = 0; in < C::FN; ++in) acc[im][in][0] = acc[im][in][1] = 0.0;

  const int wm0 = (warp % C::WARPS_M) * C::WM, wn0 = (warp / C::WARPS_M) * C::WN;
  const int fr = lane >> 2, fk = lane & 3;
  // skip warps whose whole sub-tile is out of range or strictly above the diagonal
  const bool binv = T.flags & GF_B_LINV;
  bool warp_active = (wm0 < mrows) && (wn0 < nrows);
  if ((T.flags & GF_LOWER) && (m0 + wm0 + C::WM - 1 < n0 + wn0)) warp_active = false;

  for (int kc = 0; kc < nchunks; ++kc) {
    cp_async_wait<C::STAGES - 2>();
    __syncthreads();
    const int nx = kc + C::STAGES - 1;
    if (nx < nchunks) load_chunk(nx, nx % C::STAGES);
    cp_async_commit();
    // TRSM through the inverse: B(j,k) = inv(L)(j,k) vanishes for k > j, so a warp whose columns all lie
    // below this k-chunk has nothing to add
    const bool chunk_active = warp_active && !(binv && (wn0 + C::WN <= kc * C::KC));
    if (chunk_active) {
      const double* As = smem + (kc % C::STAGES) * C::STAGE_DOUBLES;
      const double* Bs = As + C::KC * C::LDA;
#pragma unroll
      for (int k4 = 0; k4 < C::KC; k4 += 4) {
        double a[C::FM], b[C::FN];
#pragma unroll
        for (int im = 0; im < C::FM; ++im) a[im] = As[(k4 + fk) * C::LDA + wm0 + im * 8 + fr];
#pragma unroll
        for (int in = 0; in < C::FN; ++in) b[in] = Bs[(k4 + fk) * C::LDB + wn0 + in * 8 + fr];
#pragma unroll
        for (int im = 0; im < C::FM; ++im)
#pragma unroll
          for (int in = 0; in < C::FN; ++in) dmma884(acc[im][in][0], acc[im][in][1], a[im], b[in]);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  if (!warp_active) return;

  // epilogue: scatter through the relative indices
  const bool lower = T.flags & GF_LOWER, overwrite = T.flags & GF_OVERWRITE;
  double* __restrict__ Cb = lv + T.c_off;
  const int64_t ldc = T.ldc;
#pragma unroll
  for (int in = 0; in < C::FN; ++in) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = wn0 + in * 8 + fk * 2 + e;
      if (j >= nrows) continue;
      const int64_t coff = (int64_t)srel[C::TM + j] * ldc;
#pragma unroll
      for (int im = 0; im < C::FM; ++im) {
        const int i = wm0 + im * 8 + fr;
        if (i >= mrows) continue;
        if (lower && (m0 + i < n0 + j)) continue;
        double* dst = Cb + coff + srel[i];
        const double v = acc[im][in][e];
        if (overwrite) *dst = v;
        else atomicAdd(dst, -v);   // red.global.add.f64: fire-and-forget, also for exclusive targets
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small pairs: one warp per (pair, <=256-row chunk); K <= 32, N <= 32
// ------------------------------------------------------------------------------------------------
template <int KT>
__device__ __forceinline__ void small_update_rows(const GemmTask& T, int row0, int nrows, int lane, const double* Bs,
                                                  const int* srelc, double* __restrict__ lv,
                                                  const int* __restrict__ rel) {
  const double* __restrict__ src = lv + T.a_off;
  for (int i = row0 + lane; i < row0 + nrows; i += 32) {
    double a[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) a[k] = (k < T.K) ? src[(int64_t)k * T.lda + i] : 0.0;
    const int tr = rel[T.rel_off + i];
    const int jmax = min(T.N, i + 1);
    double* __restrict__ Cb = lv + T.c_off + tr;
    // two target columns at a time: the dot products are dependent FMA chains of length KT, and a rolled j loop
    // would run them one after the other
    int j = 0;
    for (; j + 1 < jmax; j += 2) {
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        d0 = fma(a[k], Bs[k * 33 + j], d0);
        d1 = fma(a[k], Bs[k * 33 + j + 1], d1);
      }
      atomicAdd(Cb + (int64_t)srelc[j] * T.ldc, -d0);
      atomicAdd(Cb + (int64_t)srelc[j + 1] * T.ldc, -d1);
    }
    if (j < jmax) {
      double dot = 0.0;
#pragma unroll
      for (int k = 0; k < KT; ++k) dot = fma(a[k], Bs[k * 33 + j], dot);
      atomicAdd(Cb + (int64_t)srelc[j] * T.ldc, -dot);
    }
  }
}

template <int KMAX>
__global__ void __launch_bounds__(128) k_update_small(const SmallTask* __restrict__ st, int count,
                                                       const GemmTask* __restrict__ tasks, double* __restrict__ lv,
                                                       const int* __restrict__ rel) {
  __shared__ double sB[4][KMAX * 33];
  __shared__ int sRelc[4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + warp;
  if (wid >= count) return;
  const SmallTask S = st[wid];
  const GemmTask T = tasks[S.pair];
  double* Bs = sB[warp];
  const double* __restrict__ src = lv + T.a_off;
  // rows K..KT of Bs are multiplied by a[k] = 0: they must hold zeros, not whatever the SM's shared memory kept
  const int KT = KMAX <= 4 ? 4 : (T.K <= 4 ? 4 : (T.K <= 8 ? 8 : (T.K <= 16 ? 16 : (T.K <= 24 ? 24 : 32))));
  // batches of 8 independent loads (a rolled loop would pay one memory latency per k)
  for (int k0 = 0; k0 < KT; k0 += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (k0 + u < T.K && lane < T.N) ? src[(int64_t)(k0 + u) * T.lda + lane] : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) if (k0 + u < KT) Bs[(k0 + u) * 33 + lane] = v[u];
  }
  if (lane < T.N) sRelc[warp][lane] = rel[T.rel_off + lane];
  __syncwarp();
  if (KMAX <= 4 || T.K <= 4) small_update_rows<4>(T, S.row0, S.nrows, lane, Bs, sRelc[warp], lv, rel);
  else if (T.K <= 8) small_update_rows<(KMAX < 8 ? KMAX : 8)>(T, S.row0, S.nrows, lane, Bs, sRelc[warp], lv, rel);
  else if (T.K <= 16) small_update_rows<(KMAX < 16 ? KMAX : 16)>(T, S.row0, S.nrows, lane, Bs, sRelc[warp], lv, rel);
  else if (T.K <= 24) small_update_rows<(KMAX < 24 ? KMAX : 24)>(T, S.row0, S.nrows, lane, Bs, sRelc[warp], lv, rel);
  else small_update_rows<KMAX>(T, S.row0, S.nrows, lane, Bs, sRelc[warp], lv, rel);
}

// ------------------------------------------------------------------------------------------------
// narrow supernodes: one warp each, POTRF (MyBLAS.h:10-25 semantics) + TRSM (MyBLAS.h:27-35) fused
// ------------------------------------------------------------------------------------------------
template <int WMAX>
__global__ void __launch_bounds__(128, WMAX > 8 ? 5 : 8) k_factor_small(const int* __restrict__ list, int count,
                                                       const SupInfo* __restrict__ sup, double* __restrict__ lv,
                                                       int* __restrict__ info) {
  constexpr int LDS = WMAX + 1;
  __shared__ double sD[4][WMAX * LDS];
  __shared__ double sR[4][WMAX];   // reciprocal diagonal
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + warp;
  if (wid >= count) return;
  const SupInfo I = sup[list[wid]];
  const int w = I.w, r = I.r;
  double* __restrict__ P = lv + I.valptr;
  double* S = sD[warp];
  double* R = sR[warp];
  // diagonal block: batches of 8 independent column loads
  for (int c0 = 0; c0 < w; c0 += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (c0 + u < w && lane < w) ? P[(int64_t)(c0 + u) * r + lane] : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) if (c0 + u < w && lane < w) S[(c0 + u) * LDS + lane] = v[u];
  }
  __syncwarp();
  for (int c = 0; c < w; ++c) {
    double acc = (lane < w) ? S[c * LDS + lane] : 0.0;
    for (int k = 0; k < c; ++k) acc = (lane < w) ? fma(-S[k * LDS + lane], S[k * LDS + c], acc) : acc;
    const double piv = __shfl_sync(0xffffffffu, acc, c);
    if (!(piv > 0.0) && lane == 0) atomicCAS(info, 0, I.col0 + c + 1);
    const double l = sqrt(piv);
    const double v = (lane == c) ? l : acc / l;
    if (lane >= c && lane < w) S[c * LDS + lane] = v;
    if (lane == c) R[c] = 1.0 / l;
    __syncwarp();
  }
  for (int c = 0; c < w; ++c)
    if (lane >= c && lane < w) P[(int64_t)c * r + lane] = S[c * LDS + lane];
  // rows below the diagonal block: x * L11' = a, one lane per row; columns in groups of 8 so that the group's loads
  // are in flight together (one memory latency per group instead of one per column)
  constexpr int G = WMAX < 8 ? WMAX : 8;
  for (int i = w + lane; i < r; i += 32) {
    double x[WMAX];
#pragma unroll
    for (int c0 = 0; c0 < WMAX; c0 += G) {
      if (c0 < w) {
#pragma unroll
        for (int u = 0; u < G; ++u) x[c0 + u] = (c0 + u < w) ? P[(int64_t)(c0 + u) * r + i] : 0.0;
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const int c = c0 + u;
          if (c < w) {
            double a = x[c];
#pragma unroll
            for (int k = 0; k < c; ++k) a = fma(-x[k], S[k * LDS + c], a);
            x[c] = a * R[c];
          }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) if (c0 + u < w) P[(int64_t)(c0 + u) * r + i] = x[c0 + u];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// block columns: POTRF of the <=128-wide diagonal block + inverse of its factor, one CTA each.
//
// Shared memory: S(i,c) at c*PLD+i (PLD = 132, conflict-free DMMA fragment loads). The block is padded with
// an identity to a multiple of 16 (nbp).  Cholesky is right-looking over 16-column macro panels, each made of
// two 8-column micro panels that every row-thread factors redundantly in registers (no intra-panel
// synchronisation: 8 rsqrt chains per micro panel), followed by a DMMA rank-16 trailing update.
// The inverse X = inv(L) is a block-row forward substitution with 16x16 blocks on DMMA:
//   X(I,J) = -D_I * sum_{K=J..I-1} L(I,K) X(K,J),  D_I = inv(L_II)
// with X(I,J), J < I, parked in the (unused) upper block (J,I) of S and the D_I in a side buffer.
// ------------------------------------------------------------------------------------------------
constexpr int PLD = 132;
constexpr int XDLD = 20;
constexpr int POTRF_THREADS = 256;
constexpr int POTRF_S = NB_MAX * PLD;                 // doubles
constexpr int POTRF_XD = (NB_MAX / 16) * 16 * XDLD;   // diagonal inverse blocks
constexpr int POTRF_T = NB_MAX * XDLD;                // per-warp 16 x 8 scratch tiles, column-major with ld XDLD
constexpr size_t POTRF_SMEM = (size_t)(POTRF_S + NB_MAX + POTRF_XD + POTRF_T) * 8;
constexpr int POTRF_LD = PLD;

// 1/sqrt(a) for the pivot chain: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22) + one third-order step — the
// scheme of the CUDA math library's rsqrt() without its range fix-ups, 4 dependent FP64 operations after the seed
// (~50 cycles against ~190 for an FP32 seed with two Newton steps, tools/fp64_latency.cu).  Non-positive pivots give
// NaN/Inf, which the caller reports through `info`.
__device__ __forceinline__ double rsqrt_pivot(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double t = a * y;
  const double e = fma(-t, y, 1.0);      // 1 - a y^2
  const double p = fma(0.375, e, 0.5);   // 1/2 + 3/8 e
  const double ye = y * e;
  return fma(ye, p, y);
}

// micro panel: columns [p0, p0+8), thread t owns row p0+t.  Every row-thread factors the 8x8 diagonal block
// redundantly in registers (no intra-panel synchronisation) with 2x2 block pivots — both reciprocal square roots of a
// column pair start together: l00 = sqrt(a), l11 = sqrt(det/a), det = a e - b^2 — and carries its own row along as a
// ninth row of the elimination.
__device__ __forceinline__ void potrf_micro8(double* S, double* rd, int p0, int nbp, int tid, int* info, int colbase,
                                             int nb) {
  const int i = p0 + tid;
  const bool active = i < nbp;
  double d[8][8], x[8];
  if (active) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int m = 0; m <= k; ++m) d[k][m] = S[(p0 + m) * PLD + p0 + k];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = S[(p0 + c) * PLD + i];
  }
  __syncthreads();   // every thread has its copy of the diagonal block before its rows are overwritten
  if (!active) return;
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    const double a = d[c][c], b = d[c + 1][c], e = d[c + 1][c + 1];
    const double det = fma(a, e, -b * b);
    if (tid == 0) {
      if (!(a > 0.0)) { if (p0 + c < nb) atomicCAS(info, 0, colbase + p0 + c + 1); }
      else if (!(det > 0.0) && p0 + c + 1 < nb) atomicCAS(info, 0, colbase + p0 + c + 2);
    }
    const double r0 = rsqrt_pivot(a);       // 1 / l00
    const double rdet = rsqrt_pivot(det);   // 1 / sqrt(det)
    const double l10 = b * r0;
    const double r1 = rdet * (a * r0);      // 1 / l11 = sqrt(a) / sqrt(det)
#pragma unroll
    for (int k = c + 2; k < 8; ++k) {
      const double x0 = d[k][c] * r0;
      d[k][c] = x0;
      d[k][c + 1] = fma(-x0, l10, d[k][c + 1]) * r1;
    }
    {
      const double x0 = x[c] * r0;
      x[c] = x0;
      x[c + 1] = fma(-x0, l10, x[c + 1]) * r1;
    }
#pragma unroll
    for (int k = c + 2; k < 8; ++k)
#pragma unroll
      for (int m = c + 2; m <= k; ++m) d[k][m] = fma(-d[k][c + 1], d[m][c + 1], fma(-d[k][c], d[m][c], d[k][m]));
#pragma unroll
    for (int m = c + 2; m < 8; ++m) x[m] = fma(-x[c + 1], d[m][c + 1], fma(-x[c], d[m][c], x[m]));
    if (tid == 0) { rd[p0 + c] = r0; rd[p0 + c + 1] = r1; }
  }
  // a row inside the diagonal block reproduces L_dd itself up to its diagonal entry; what lies right of it is scratch
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (p0 + c <= i) S[(p0 + c) * PLD + i] = x[c];
}

// rank-8 update of columns [p0+8, p0+16) by the micro panel [p0, p0+8) on DMMA: 8x8 tiles of rows, K = 8
__device__ __forceinline__ void potrf_mid8(double* S, int p0, int nbp, int warp, int lane) {
  const int fr = lane >> 2, fk = lane & 3;
  const int q0 = p0 + 8;
  const int nt = (nbp - q0) >> 3;
  const double b0 = S[(p0 + fk) * PLD + q0 + fr], b1 = S[(p0 + 4 + fk) * PLD + q0 + fr];
  for (int t = warp; t < nt; t += POTRF_THREADS / 32) {
    const int r0 = q0 + t * 8;
    double c0 = S[(q0 + fk * 2) * PLD + r0 + fr], c1 = S[(q0 + fk * 2 + 1) * PLD + r0 + fr];
    const double a0 = -S[(p0 + fk) * PLD + r0 + fr], a1 = -S[(p0 + 4 + fk) * PLD + r0 + fr];
    dmma884(c0, c1, a0, b0);
    dmma884(c0, c1, a1, b1);
    S[(q0 + fk * 2) * PLD + r0 + fr] = c0;
    S[(q0 + fk * 2 + 1) * PLD + r0 + fr] = c1;
  }
}

// DMMA trailing update: S(i,k) -= sum_{c<16} S(i,p0+c) S(k,p0+c) for i >= k >= q0 = p0+16, 16x16 tiles per warp
__device__ __forceinline__ void potrf_trailing16(double* S, int p0, int nbp, int warp, int lane) {
  const int q0 = p0 + 16;
  const int T16 = (nbp - q0) >> 4;
  const int ntiles = T16 * (T16 + 1) / 2;
  const int fr = lane >> 2, fk = lane & 3;
  for (int tile = warp; tile < ntiles; tile += POTRF_THREADS / 32) {
    int ti = 0, rem = tile;
    while (rem > ti) { rem -= ti + 1; ++ti; }
    const int tj = rem;
    const int r0 = q0 + ti * 16, c0 = q0 + tj * 16;
    const bool diag = ti == tj;
    double c[2][2][2];
#pragma unroll
    for (int fi = 0; fi < 2; ++fi)
#pragma unroll
      for (int fj = 0; fj < 2; ++fj)
#pragma unroll
        for (int e = 0; e < 2; ++e) c[fi][fj][e] = S[(c0 + fj * 8 + fk * 2 + e) * PLD + r0 + fi * 8 + fr];
#pragma unroll
    for (int k4 = 0; k4 < 16; k4 += 4) {
      double a[2], b[2];
#pragma unroll
      for (int fi = 0; fi < 2; ++fi) a[fi] = -S[(p0 + k4 + fk) * PLD + r0 + fi * 8 + fr];
#pragma unroll
      for (int fj = 0; fj < 2; ++fj) b[fj] = S[(p0 + k4 + fk) * PLD + c0 + fj * 8 + fr];
#pragma unroll
      for (int fi = 0; fi < 2; ++fi)
#pragma unroll
        for (int fj = 0; fj < 2; ++fj) dmma884(c[fi][fj][0], c[fi][fj][1], a[fi], b[fj]);
    }
#pragma unroll
    for (int fi = 0; fi < 2; ++fi)
#pragma unroll
      for (int fj = 0; fj < 2; ++fj) {
        if (diag && fi == 0 && fj == 1) continue;   // strictly above the diagonal
#pragma unroll
        for (int e = 0; e < 2; ++e) S[(c0 + fj * 8 + fk * 2 + e) * PLD + r0 + fi * 8 + fr] = c[fi][fj][e];
      }
  }
}

// finished columns [p0, p0+16) of the block -> global panel (rows >= column, below nb); fire-and-forget stores that
// overlap the trailing update instead of a write-back pass at the end
__device__ __forceinline__ void store_panel16(const double* S, double* __restrict__ P, int64_t r, int nb, int p0, int tid) {
  const int i = p0 + (tid & 127);
  if (i >= nb) return;
  const int cb = p0 + (tid >> 7) * 8;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int c = cb + u;
    if (c <= i) P[(int64_t)c * r + i] = S[c * PLD + i];
  }
}

// Cholesky of the padded nbp x nbp block held in S (lower part), written to the global panel P (ld r) as the
// columns finish; fills rd = 1/diag(L).  Right-looking over 16-column macro panels, each made of two 8-column micro
// panels joined by a DMMA rank-8 update, followed by a DMMA rank-16 trailing update.
__device__ __forceinline__ void potrf_in_smem(double* S, double* rd, int nb, int nbp, int tid, int* info, int colbase,
                                              double* __restrict__ P, int64_t r) {
  const int warp = tid >> 5, lane = tid & 31;
  for (int p0 = 0; p0 < nbp; p0 += 16) {
    potrf_micro8(S, rd, p0, nbp, tid, info, colbase, nb);
    __syncthreads();
    potrf_mid8(S, p0, nbp, warp, lane);
    __syncthreads();
    potrf_micro8(S, rd, p0 + 8, nbp, tid, info, colbase, nb);
    __syncthreads();
    if (P) store_panel16(S, P, r, nb, p0, tid);
    if (p0 + 16 < nbp) {
      potrf_trailing16(S, p0, nbp, warp, lane);
      __syncthreads();
    }
  }
}

// X = inv(L) -> global Xg (column-major, ld NB_MAX, lower part; the rest of the slot stays zero).
//   D_I = inv(L_II) for the 16x16 diagonal blocks (16 threads per block, one column each), kept in XD;
//   then every 8-column group g of X is an independent block forward substitution down its block rows,
//     X(I,g) = -D_I * sum_{K=J..I-1} L(I,K) X(K,g),     J = block of g,
//   run by one warp on DMMA with only warp-level synchronisation: the group's finished blocks are parked in the
//   unused upper block (J,I) of S (read back as B operands by the same warp) and stored to global straight from the
//   accumulator fragments.  Groups g and 15-g share a warp (long chains with short ones).
__device__ __forceinline__ void invert_in_smem(double* S, const double* rd, double* XD, double* Tt, double* __restrict__ Xg,
                                               int nb, int nbp, int tid) {
  const int warp = tid >> 5, lane = tid & 31, fr = lane >> 2, fk = lane & 3;
  const int nI = nbp >> 4;
  if (tid < nI * 16) {
    const int I = tid >> 4, c = tid & 15, o = I * 16;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double v = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) v = fma(-S[(o + k) * PLD + o + i], x[k], v);
      x[i] = v * rd[o + i];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) XD[I * 16 * XDLD + c * XDLD + i] = x[i];   // zero above the diagonal (i < c)
  }
  __syncthreads();
  for (int idx = tid; idx < nI * 256; idx += POTRF_THREADS) {
    const int I = idx >> 8, c = (idx >> 4) & 15, i = idx & 15;
    const int row = I * 16 + i, col = I * 16 + c;
    if (row < nb && col < nb) Xg[col * NB_MAX + row] = XD[I * 16 * XDLD + c * XDLD + i];
  }
  double* T = Tt + warp * 8 * XDLD;   // warp-private 16 x 8 scratch
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int g = half == 0 ? warp : 15 - warp;
    const int J = g >> 1, cj = (g & 1) * 8;
#pragma unroll 1
    for (int I = J + 1; I < nI; ++I) {
      // T = sum_K L(I,K) X(K,g): two row fragments, each split over two accumulator pairs (even / odd k4 steps)
      double t[2][2][2];
#pragma unroll
      for (int fi = 0; fi < 2; ++fi)
#pragma unroll
        for (int h = 0; h < 2; ++h) t[fi][h][0] = t[fi][h][1] = 0.0;
#pragma unroll 1
      for (int K = J; K < I; ++K) {
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
          const double b = (K == J) ? XD[J * 16 * XDLD + (cj + fr) * XDLD + k4 + fk]
                                    : S[(K * 16 + cj + fr) * PLD + J * 16 + k4 + fk];
#pragma unroll
          for (int fi = 0; fi < 2; ++fi) {
            const double a = S[(K * 16 + k4 + fk) * PLD + I * 16 + fi * 8 + fr];
            dmma884(t[fi][(k4 >> 2) & 1][0], t[fi][(k4 >> 2) & 1][1], a, b);
          }
        }
      }
#pragma unroll
      for (int fi = 0; fi < 2; ++fi) {
        T[(fk * 2) * XDLD + fi * 8 + fr] = t[fi][0][0] + t[fi][1][0];
        T[(fk * 2 + 1) * XDLD + fi * 8 + fr] = t[fi][0][1] + t[fi][1][1];
      }
      __syncwarp();
      // X(I,g) = -D_I T
      double xv[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int k4 = 0; k4 < 16; k4 += 4) {
        const double b = T[fr * XDLD + k4 + fk];
#pragma unroll
        for (int fi = 0; fi < 2; ++fi) {
          const double a = -XD[I * 16 * XDLD + (k4 + fk) * XDLD + fi * 8 + fr];
          dmma884(xv[fi][0], xv[fi][1], a, b);
        }
      }
#pragma unroll
      for (int fi = 0; fi < 2; ++fi)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int lr = fi * 8 + fr, lc = cj + fk * 2 + e;
          S[(I * 16 + lc) * PLD + J * 16 + lr] = xv[fi][e];
          const int row = I * 16 + lr, col = J * 16 + lc;
          if (row < nb && col < nb) Xg[col * NB_MAX + row] = xv[fi][e];
        }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ void load_padded_block(double* S, const double* __restrict__ P, int64_t r, int nb, int nbp,
                                                  int tid) {
  // row = tid & 127, two column phases; batches of 16 independent global loads before the dependent shared stores
  constexpr int U = 16;
  const int i = tid & 127;
  if (i >= nbp) return;
  for (int c0 = tid >> 7; c0 <= i; c0 += 2 * U) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + 2 * u;
      v[u] = (c <= i && i < nb && c < nb) ? P[(int64_t)c * r + i] : ((i == c) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + 2 * u;
      if (c <= i) S[c * PLD + i] = v[u];
    }
  }
}

// smem_cols = widest (padded) block of the launch: narrower steps take less shared memory and co-reside on an SM
__host__ __device__ constexpr size_t potrf_smem_bytes(int smem_cols) {
  return (size_t)(smem_cols * PLD + NB_MAX + POTRF_XD + POTRF_T) * 8;
}
__global__ void __launch_bounds__(POTRF_THREADS) k_potrf_block(const BlockTask* __restrict__ bt,
                                                                const SupInfo* __restrict__ sup,
                                                                double* __restrict__ lv, double* __restrict__ linv,
                                                                int* __restrict__ info, int smem_cols) {
  extern __shared__ __align__(16) double smem[];
  double* S = smem;
  double* rd = S + smem_cols * PLD;
  double* XD = rd + NB_MAX;
  double* Tt = XD + POTRF_XD;
  const BlockTask B = bt[blockIdx.x];
  const SupInfo I = sup[B.sup];
  const int nb = B.nb, nbp = (nb + 15) & ~15, tid = threadIdx.x;
  const int64_t r = I.r;
  double* __restrict__ P = lv + I.valptr + (int64_t)B.j0 * r + B.j0;   // (j0, j0) of the panel
  load_padded_block(S, P, r, nb, nbp, tid);
  __syncthreads();
  potrf_in_smem(S, rd, nb, nbp, tid, info, I.col0 + B.j0, P, r);
  invert_in_smem(S, rd, XD, Tt, linv + (int64_t)B.slot * NB_MAX * NB_MAX, nb, nbp, tid);
}

// inverse diagonal blocks for a factor that was produced elsewhere (parsy_cuda_set_factor, drop-in solves)
__global__ void __launch_bounds__(POTRF_THREADS) k_invert_block(const BlockTask* __restrict__ bt,
                                                                 const SupInfo* __restrict__ sup,
                                                                 const double* __restrict__ lv,
                                                                 double* __restrict__ linv) {
  extern __shared__ __align__(16) double smem[];
  double* S = smem;
  double* rd = S + POTRF_S;
  double* XD = rd + NB_MAX;
  double* Tt = XD + POTRF_XD;
  const BlockTask B = bt[blockIdx.x];
  const SupInfo I = sup[B.sup];
  const int nb = B.nb, nbp = (nb + 15) & ~15, tid = threadIdx.x;
  const int64_t r = I.r;
  const double* __restrict__ P = lv + I.valptr + (int64_t)B.j0 * r + B.j0;
  load_padded_block(S, P, r, nb, nbp, tid);
  __syncthreads();
  if (tid < nbp) rd[tid] = 1.0 / S[tid * PLD + tid];
  __syncthreads();
  invert_in_smem(S, rd, XD, Tt, linv + (int64_t)B.slot * NB_MAX * NB_MAX, nb, nbp, tid);
}

// ------------------------------------------------------------------------------------------------
// structure-time helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_row(const int* __restrict__ rows, int len, int key) {
  int lo = 0, hi = len - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rows[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// rel[e]: local row (in the target supernode's row list) of every row lb.. of the descendant, per pair
__global__ void k_build_rel(int64_t total, int npairs, const int64_t* __restrict__ prefix,
                            const int* __restrict__ psrc, const int* __restrict__ ptgt, const int* __restrict__ plb,
                            const SupInfo* __restrict__ sup, const int* __restrict__ lR, int* __restrict__ rel) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = npairs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= e) lo = mid; else hi = mid - 1;
    }
    const int i = (int)(e - prefix[lo]);
    const SupInfo D = sup[psrc[lo]], T = sup[ptgt[lo]];
    const int row = lR[D.rowptr + plb[lo] + i];
    int v;
    if (row < T.col0 + T.w) v = row - T.col0;
    else v = T.w + find_row(lR + T.rowptr + T.w, T.r - T.w, row);
    rel[e] = v;
  }
}

// a_pos[p]: offset in lValues of A entry p (column j, row r[p]) — the `map` of parallel_PB_Cholesky_05.h:100-112
// skip[s] != 0: this handle does not assemble supernode s (sharded plans) — position -1
__global__ void k_build_apos(int64_t nnz, int n, const int* __restrict__ c, const int* __restrict__ r,
                             const int* __restrict__ col2sup, const SupInfo* __restrict__ sup,
                             const int* __restrict__ lR, const unsigned char* __restrict__ skip,
                             int64_t* __restrict__ apos) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (c[mid] <= p) lo = mid; else hi = mid - 1;
    }
    const int j = lo;
    if (skip && skip[col2sup[j]]) { apos[p] = -1; continue; }
    const SupInfo I = sup[col2sup[j]];
    const int row = r[p];
    int pos;
    if (row < I.col0 + I.w) pos = row - I.col0;
    else pos = I.w + find_row(lR + I.rowptr + I.w, I.r - I.w, row);
    apos[p] = I.valptr + (int64_t)(j - I.col0) * I.r + pos;
  }
}

__global__ void k_assemble(int64_t nnz, const int64_t* __restrict__ apos, const double* __restrict__ vals,
                           double* __restrict__ lv) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = apos[p];
    if (q >= 0) lv[q] = vals[p];
  }
}

// dst[i] += sum over the other buffers (single-process emulation of the ranks' all-reduce; tests only)
__global__ void k_sum_buffers(int64_t count, double* __restrict__ dst, const double* const* __restrict__ src, int nsrc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    double a = 0.0;
    for (int k = 0; k < nsrc; ++k) a += src[k][i];
    dst[i] = a;
  }
}
// timeline probes (parsy_cuda_sharded_trace_top): a one-thread kernel node that stores the GPU's nanosecond timer
__global__ void k_stamp(unsigned long long* out) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  *out = t;
}
__global__ void k_zero_range(double* __restrict__ x, int b, int e) {
  for (int i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x) x[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// triangular sweeps
// ------------------------------------------------------------------------------------------------
// forward, narrow supernodes: x_s = L11^-1 y_s ; y[rows below] -= L21 x_s          (Triangular_BCSC.h:206-224)
__global__ void __launch_bounds__(128) k_fwd_small(const int* __restrict__ list, int count,
                                                    const SupInfo* __restrict__ sup, const int* __restrict__ lR,
                                                    const double* __restrict__ lv, double* __restrict__ y,
                                                    double* __restrict__ xs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + warp;
  if (wid >= count) return;
  const SupInfo I = sup[list[wid]];
  const int w = I.w, r = I.r;
  const double* __restrict__ P = lv + I.valptr;
  double xv = (lane < w) ? y[I.col0 + lane] : 0.0;
  for (int c = 0; c < w; ++c) {
    const double lcol = (lane >= c && lane < w) ? P[(int64_t)c * r + lane] : 1.0;   // column c of L11
    const double xc = __shfl_sync(0xffffffffu, xv, c) / __shfl_sync(0xffffffffu, lcol, c);
    if (lane == c) xv = xc;
    else if (lane > c && lane < w) xv = fma(-lcol, xc, xv);
  }
  if (lane < w) xs[I.col0 + lane] = xv;
  const int* __restrict__ rows = lR + I.rowptr;
  for (int i0 = w; i0 < r; i0 += 32) {
    const int i = i0 + lane;
    double t = 0.0;
    for (int c = 0; c < w; ++c) {
      const double xc = __shfl_sync(0xffffffffu, xv, c);
      if (i < r) t = fma(P[(int64_t)c * r + i], xc, t);
    }
    if (i < r) atomicAdd(&y[rows[i]], -t);
  }
}

// forward, block columns: every CTA recomputes x_b = inv(L_bb) y_b (nb x nb GEMV from the inverse store),
// tile 0 publishes it to xs, and each CTA applies its 64-row slice of L21 to y.
constexpr int SOLVE_ROWS = 64;
__global__ void __launch_bounds__(256) k_fwd_block(const BlockTask* __restrict__ bt, int ntasks,
                                                    const SupInfo* __restrict__ sup, const int* __restrict__ lR,
                                                    const double* __restrict__ lv, const double* __restrict__ linv,
                                                    double* __restrict__ y, double* __restrict__ xs) {
  __shared__ double sy[NB_MAX], sx[NB_MAX], spart[4][SOLVE_ROWS];
  const int tid = threadIdx.x, bid = blockIdx.x;
  int lo = 0, hi = ntasks - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (bt[mid].tile0 <= bid) lo = mid; else hi = mid - 1;
  }
  const BlockTask B = bt[lo];
  const SupInfo I = sup[B.sup];
  const int tile = bid - B.tile0, nb = B.nb, r = I.r;
  const int cbase = I.col0 + B.j0;
  if (tid < nb) sy[tid] = y[cbase + tid];
  __syncthreads();
  const double* __restrict__ X = linv + (int64_t)B.slot * NB_MAX * NB_MAX;
  if (tid < nb) {
    double acc = 0.0;
    for (int k = 0; k <= tid; ++k) acc = fma(X[k * NB_MAX + tid], sy[k], acc);
    sx[tid] = acc;
    if (tile == 0) xs[cbase + tid] = acc;
  }
  __syncthreads();
  const int rbeg = B.j0 + nb + tile * SOLVE_ROWS;       // local row in the supernode
  const int nr = min(SOLVE_ROWS, r - rbeg);
  if (nr <= 0) return;
  const int ri = tid & 63, cg = tid >> 6;
  const double* __restrict__ P = lv + I.valptr + (int64_t)B.j0 * r + rbeg;
  double t = 0.0;
  if (ri < nr)
    for (int c = cg; c < nb; c += 4) t = fma(P[(int64_t)c * r + ri], sx[c], t);
  spart[cg][ri] = t;
  __syncthreads();
  if (tid < nr) {
    const double s = spart[0][tid] + spart[1][tid] + spart[2][tid] + spart[3][tid];
    atomicAdd(&y[lR[I.rowptr + rbeg + tid]], -s);
  }
}

// backward, narrow supernodes: x_s = L11^-T (y_s - L21' x[rows below])
__global__ void __launch_bounds__(128) k_bwd_small(const int* __restrict__ list, int count,
                                                    const SupInfo* __restrict__ sup, const int* __restrict__ lR,
                                                    const double* __restrict__ lv, double* __restrict__ x) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + warp;
  if (wid >= count) return;
  const SupInfo I = sup[list[wid]];
  const int w = I.w, r = I.r;
  const double* __restrict__ P = lv + I.valptr;
  const int* __restrict__ rows = lR + I.rowptr;
  double mine = (lane < w) ? x[I.col0 + lane] : 0.0;   // lane c holds y_c
  // t_c = sum_i L(i,c) x[rows[i]]: lanes stride the rows, shuffle-reduce per column
  for (int c = 0; c < w; ++c) {
    double part = 0.0;
    for (int i = w + lane; i < r; i += 32) part = fma(P[(int64_t)c * r + i], x[rows[i]], part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == c) mine -= part;
  }
  // L11' x = mine, backward substitution; lane j holds x_j
  for (int c = w - 1; c >= 0; --c) {
    const double lcol = (lane >= c && lane < w) ? P[(int64_t)c * r + lane] : 0.0;   // L(lane, c)
    double part = (lane > c && lane < w) ? lcol * mine : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const double d = __shfl_sync(0xffffffffu, lcol, c);
    if (lane == c) mine = (mine - part) / d;
  }
  if (lane < w) x[I.col0 + lane] = mine;
}

// backward, block columns, phase A: x_b -= L21_b' x[rows below] (64-row slices, atomics into x_b)
__global__ void __launch_bounds__(256) k_bwd_block_gemv(const BlockTask* __restrict__ bt, int ntasks,
                                                         const SupInfo* __restrict__ sup, const int* __restrict__ lR,
                                                         const double* __restrict__ lv, double* __restrict__ x) {
  __shared__ double sx[SOLVE_ROWS];
  const int tid = threadIdx.x, bid = blockIdx.x, lane = tid & 31, warp = tid >> 5;
  int lo = 0, hi = ntasks - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (bt[mid].tile0 <= bid) lo = mid; else hi = mid - 1;
  }
  const BlockTask B = bt[lo];
  const SupInfo I = sup[B.sup];
  const int tile = bid - B.tile0, nb = B.nb, r = I.r;
  const int rbeg = B.j0 + nb + tile * SOLVE_ROWS;
  const int nr = min(SOLVE_ROWS, r - rbeg);
  if (nr <= 0) return;
  if (tid < nr) sx[tid] = x[lR[I.rowptr + rbeg + tid]];
  __syncthreads();
  const double* __restrict__ P = lv + I.valptr + (int64_t)B.j0 * r + rbeg;
  // each warp takes columns warp, warp+8, ...; lanes stride the 64 rows; shuffle-reduce
  for (int c = warp; c < nb; c += 8) {
    double part = 0.0;
    for (int i = lane; i < nr; i += 32) part = fma(P[(int64_t)c * r + i], sx[i], part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) atomicAdd(&x[I.col0 + B.j0 + c], -part);
  }
}

// backward, block columns, phase B: x_b = inv(L_bb)' x_b
__global__ void __launch_bounds__(128) k_bwd_block_diag(const BlockTask* __restrict__ bt,
                                                         const SupInfo* __restrict__ sup,
                                                         const double* __restrict__ linv, double* __restrict__ x) {
  __shared__ double sy[NB_MAX];
  const BlockTask B = bt[blockIdx.x];
  const SupInfo I = sup[B.sup];
  const int tid = threadIdx.x, nb = B.nb, cbase = I.col0 + B.j0;
  if (tid < nb) sy[tid] = x[cbase + tid];
  __syncthreads();
  const double* __restrict__ X = linv + (int64_t)B.slot * NB_MAX * NB_MAX;
  if (tid < nb) {
    // (X' y)_c = sum_{k >= c} X(k,c) y_k ; column c of X is contiguous
    double acc = 0.0;
    for (int k = tid; k < nb; ++k) acc = fma(X[tid * NB_MAX + k], sy[k], acc);
    x[cbase + tid] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// single-launch dataflow sweeps
//
// One launch per sweep. CTAs take tickets (atomic counter) and walk the solve tasks in dependency order, so every
// producer of a task holds a smaller ticket and is already running: spinning on its counter cannot deadlock.
// forward : node ready  <=> done[node] == need[node]  (every task that adds into its right-hand side has finished)
// backward: task ready  <=> solved[u] != 0 for every node u its rows belong to;  the last slice of a block column
//           to finish its partial L21' x does the diagonal solve and publishes solved[node].
// Right-hand-side entries are only ever modified by L2 atomics and read with ld.global.cg after an acquire.
// Everything that does not depend on the right-hand side (the slice of L, the diagonal block of a narrow
// supernode) is pulled into registers BEFORE the task waits for its inputs: the dependency chain then only
// carries the arithmetic, not the HBM latency.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
__device__ __forceinline__ void spin_until_ge_busy(const int* p, int target) {
  while (ld_acquire_gpu(p) < target) {}
}

constexpr int SWEEP_THREADS = 256;
constexpr size_t FWD_SWEEP_SMEM = (size_t)(NB_MAX * (NB_MAX + 1) / 2) * 8;   // packed lower triangle of one inverse block

// ------------------------------------------------------------------------------------------------
// Narrow supernodes (width <= SMALL_W) in the sweeps.  Two task shapes, shared by the general dataflow kernels and
// the light narrow-only kernels below:
//   * warp task   — one warp per supernode, eight per CTA (SolveCta kind 0);
//   * tall task   — one CTA per supernode whose panel below the diagonal block is long (kind 2): warp 0 solves the
//                   diagonal block, all eight warps stream the rows (256 per pass).  A single warp would walk such a
//                   panel 32 rows at a time, ~1 us per pass, right on the dependency chain (measured with
//                   parsy_cuda_sweep_trace: up to 114 us per task on the 2-D 1000x1000 grid).
// The diagonal block sits in shared memory (packed lower triangle, column c at c*w - c(c-1)/2) with the reciprocal
// diagonal beside it; it, the row indices and the first columns of the first pass are fetched BEFORE the task waits for
// its inputs, and panel loads are issued in batches so that one memory latency covers several columns.
// ------------------------------------------------------------------------------------------------
constexpr int NARROW_TRI = SMALL_W * (SMALL_W + 1) / 2;
constexpr int NARROW_PF = 4;        // columns of the first 32 rows a warp task prefetches
constexpr int TALL_PF = 8;          // columns of the first 256 rows a tall task prefetches
constexpr int NARROW_SCRATCH = NARROW_TRI + SMALL_W;   // doubles of shared memory per warp task: L, R

__device__ __forceinline__ void narrow_stage_diag(double* L, double* R, const double* __restrict__ P, int w, int r, int lane) {
  for (int c0 = 0; c0 < w; c0 += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + u;
      v[u] = (c < w && lane >= c && lane < w) ? P[(int64_t)c * r + lane] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + u;
      if (c < w && lane >= c && lane < w) L[c * w - (c * (c - 1)) / 2 + (lane - c)] = v[u];
    }
  }
  __syncwarp();
  if (lane < w) R[lane] = 1.0 / L[lane * w - (lane * (lane - 1)) / 2];
  __syncwarp();
}
// x_b = inv(L11) y_b by one warp (lane = row), L11 and 1/diag in shared memory
__device__ __forceinline__ double narrow_fwd_subst(double xv, const double* L, const double* R, int w, int lane) {
  for (int c = 0; c < w; ++c) {
    const double xc = __shfl_sync(0xffffffffu, xv, c) * R[c];
    if (lane == c) xv = xc;
    else if (lane > c && lane < w) xv = fma(-L[c * w - (c * (c - 1)) / 2 + (lane - c)], xc, xv);
  }
  return xv;
}
// x_b = inv(L11') t, column-oriented: x_c is final once every x_k, k > c, has been eliminated from it
__device__ __forceinline__ double narrow_bwd_subst(double mine, const double* L, const double* R, int w, int lane) {
  for (int c = w - 1; c >= 0; --c) {
    const double xc = __shfl_sync(0xffffffffu, mine, c) * R[c];
    if (lane == c) mine = xc;
    else if (lane < c) mine = fma(-L[lane * w - (lane * (lane - 1)) / 2 + (c - lane)], xc, mine);
  }
  return mine;
}

__device__ __forceinline__ SupInfo narrow_info(const SolveTask& T) {
  SupInfo I;
  I.rowptr = T.rowptr; I.valptr = T.valptr; I.col0 = T.col0; I.w = T.nb; I.r = T.r; I.flags = 1;
  return I;
}

__device__ __forceinline__ void fwd_narrow_warp_task(const SolveTask& T, const SupInfo& I, double* L, double* R, int lane,
                                                     const int* __restrict__ targets, const int* __restrict__ need,
                                                     int* __restrict__ done, const int* __restrict__ lR,
                                                     const double* __restrict__ lv, double* __restrict__ y,
                                                     double* __restrict__ xs) {
  const int w = I.w, r = I.r;
  const double* __restrict__ P = lv + I.valptr;
  const int* __restrict__ rows = lR + I.rowptr;
  const int ifirst = w + lane;
  const int row_first = (ifirst < r) ? rows[ifirst] : -1;
  double pf[NARROW_PF];
#pragma unroll
  for (int u = 0; u < NARROW_PF; ++u) pf[u] = (u < w && ifirst < r) ? P[(int64_t)u * r + ifirst] : 0.0;
  const int tq = (T.tgt_begin + lane < T.tgt_end) ? targets[T.tgt_begin + lane] : -1;
  narrow_stage_diag(L, R, P, w, r, lane);
  if (lane == 0) spin_until_ge_busy(&done[T.node], T.need);
  __syncwarp();
  double xv = (lane < w) ? __ldcg(&y[I.col0 + lane]) : 0.0;
  xv = narrow_fwd_subst(xv, L, R, w, lane);
  if (lane < w) xs[I.col0 + lane] = xv;
  for (int i0 = w; i0 < r; i0 += 32) {
    const int i = i0 + lane;
    const bool firstc = i0 == w;
    double t = 0.0;
    int c = 0;
    if (firstc) {
#pragma unroll
      for (int u = 0; u < NARROW_PF; ++u)
        if (u < w) t = fma(pf[u], __shfl_sync(0xffffffffu, xv, u), t);
      c = min(w, NARROW_PF);
    }
    for (; c < w; c += 4) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (c + u < w && i < r) ? P[(int64_t)(c + u) * r + i] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) t = fma(v[u], __shfl_sync(0xffffffffu, xv, (c + u) & 31), t);
    }
    const int row = firstc ? row_first : (i < r ? rows[i] : -1);
    if (row >= 0) atomicAdd(&y[row], -t);
  }
  __threadfence();
  __syncwarp();
  if (tq >= 0) atomicAdd(&done[tq], 1);
  for (int q = T.tgt_begin + 32 + lane; q < T.tgt_end; q += 32) atomicAdd(&done[targets[q]], 1);
}

// scratch: L[NARROW_TRI] R[SMALL_W] sxv[SMALL_W]
__device__ __forceinline__ void fwd_narrow_tall_task(const SolveTask& T, const SupInfo& I, double* scratch, int tid,
                                                     const int* __restrict__ targets, const int* __restrict__ need,
                                                     int* __restrict__ done, const int* __restrict__ lR,
                                                     const double* __restrict__ lv, double* __restrict__ y,
                                                     double* __restrict__ xs) {
  double* L = scratch;
  double* R = scratch + NARROW_TRI;
  double* sxv = R + SMALL_W;
  const int lane = tid & 31, warp = tid >> 5;
  const int w = I.w, r = I.r;
  const double* __restrict__ P = lv + I.valptr;
  const int* __restrict__ rows = lR + I.rowptr;
  const int ifirst = w + tid;
  const int row_first = (ifirst < r) ? rows[ifirst] : -1;
  double pf[TALL_PF];
#pragma unroll
  for (int u = 0; u < TALL_PF; ++u) pf[u] = (u < w && ifirst < r) ? P[(int64_t)u * r + ifirst] : 0.0;
  const int tq = (T.tgt_begin + tid < T.tgt_end) ? targets[T.tgt_begin + tid] : -1;
  if (warp == 0) narrow_stage_diag(L, R, P, w, r, lane);
  if (tid == 0) spin_until_ge_busy(&done[T.node], T.need);
  __syncthreads();
  if (warp == 0) {
    double xv = (lane < w) ? __ldcg(&y[I.col0 + lane]) : 0.0;
    xv = narrow_fwd_subst(xv, L, R, w, lane);
    if (lane < w) xs[I.col0 + lane] = xv;
    sxv[lane] = (lane < w) ? xv : 0.0;
  }
  __syncthreads();
  for (int i0 = w; i0 < r; i0 += SWEEP_THREADS) {
    const int i = i0 + tid;
    const bool firstc = i0 == w;
    double t = 0.0;
    int c = 0;
    if (firstc) {
#pragma unroll
      for (int u = 0; u < TALL_PF; ++u) t = fma(pf[u], sxv[u], t);
      c = min(w, TALL_PF);
    }
    for (; c < w; c += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (c + u < w && i < r) ? P[(int64_t)(c + u) * r + i] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) t = fma(v[u], sxv[(c + u) & (SMALL_W - 1)], t);
    }
    const int row = firstc ? row_first : (i < r ? rows[i] : -1);
    if (row >= 0) atomicAdd(&y[row], -t);
  }
  __threadfence();
  __syncthreads();
  if (tq >= 0) atomicAdd(&done[tq], 1);
  for (int q = T.tgt_begin + SWEEP_THREADS + tid; q < T.tgt_end; q += SWEEP_THREADS) atomicAdd(&done[targets[q]], 1);
}

__device__ __forceinline__ void bwd_narrow_warp_task(const SolveTask& T, const SupInfo& I, double* L, double* R, int lane,
                                                     const int* __restrict__ targets, int* __restrict__ solved,
                                                     const int* __restrict__ lR, const double* __restrict__ lv,
                                                     double* __restrict__ x) {
  const int w = I.w, r = I.r;
  const double* __restrict__ P = lv + I.valptr;
  const int* __restrict__ rows = lR + I.rowptr;
  const int ifirst = w + lane;
  const int row_first = (ifirst < r) ? rows[ifirst] : -1;
  double pf[NARROW_PF];
#pragma unroll
  for (int u = 0; u < NARROW_PF; ++u) pf[u] = (u < w && ifirst < r) ? P[(int64_t)u * r + ifirst] : 0.0;
  narrow_stage_diag(L, R, P, w, r, lane);
  for (int q = T.tgt_begin + lane; q < T.tgt_end; q += 32) spin_until_ge_busy(&solved[targets[q]], 1);
  __syncwarp();
  double mine = (lane < w) ? __ldcg(&x[I.col0 + lane]) : 0.0;
  const double xr_first = (row_first >= 0) ? __ldcg(&x[row_first]) : 0.0;
  // t_c = sum_i L(i,c) x[rows[i]], four columns at a time: per-lane partial sums over the lane's rows, then one
  // interleaved shuffle reduction per group
  for (int c0 = 0; c0 < w; c0 += 4) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i0 = w; i0 < r; i0 += 32) {
      const int ii = i0 + lane;
      if (i0 == w) {
        if (c0 == 0) {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fma(pf[u], xr_first, acc[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = (c0 + u < w && ii < r) ? fma(P[(int64_t)(c0 + u) * r + ii], xr_first, acc[u]) : acc[u];
        }
      } else {
        const double xr = (ii < r) ? __ldcg(&x[rows[ii]]) : 0.0;
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (c0 + u < w && ii < r) ? P[(int64_t)(c0 + u) * r + ii] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fma(v[u], xr, acc[u]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) if (lane == c0 + u) mine -= acc[u];
  }
  mine = narrow_bwd_subst(mine, L, R, w, lane);
  if (lane < w) x[I.col0 + lane] = mine;
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicExch(&solved[T.node], 1);
}

// scratch: L[NARROW_TRI] R[SMALL_W] sred[8 warps][8 columns]
__device__ __forceinline__ void bwd_narrow_tall_task(const SolveTask& T, const SupInfo& I, double* scratch, int tid,
                                                     const int* __restrict__ targets, int* __restrict__ solved,
                                                     const int* __restrict__ lR, const double* __restrict__ lv,
                                                     double* __restrict__ x) {
  double* L = scratch;
  double* R = scratch + NARROW_TRI;
  double* sred = R + SMALL_W;
  const int lane = tid & 31, warp = tid >> 5;
  const int w = I.w, r = I.r;
  const double* __restrict__ P = lv + I.valptr;
  const int* __restrict__ rows = lR + I.rowptr;
  const int ifirst = w + tid;
  const int row_first = (ifirst < r) ? rows[ifirst] : -1;
  double pf[TALL_PF];
#pragma unroll
  for (int u = 0; u < TALL_PF; ++u) pf[u] = (u < w && ifirst < r) ? P[(int64_t)u * r + ifirst] : 0.0;
  if (warp == 0) narrow_stage_diag(L, R, P, w, r, lane);
  for (int q = T.tgt_begin + tid; q < T.tgt_end; q += SWEEP_THREADS) spin_until_ge_busy(&solved[targets[q]], 1);
  __syncthreads();
  double mine = (warp == 0 && lane < w) ? __ldcg(&x[I.col0 + lane]) : 0.0;
  const double xr_first = (row_first >= 0) ? __ldcg(&x[row_first]) : 0.0;
  for (int c0 = 0; c0 < w; c0 += 8) {
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.0;
    for (int i0 = w; i0 < r; i0 += SWEEP_THREADS) {
      const int ii = i0 + tid;
      if (i0 == w && c0 == 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fma(pf[u], xr_first, acc[u]);
      } else {
        const double xr = (i0 == w) ? xr_first : ((ii < r) ? __ldcg(&x[rows[ii]]) : 0.0);
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (c0 + u < w && ii < r) ? P[(int64_t)(c0 + u) * r + ii] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fma(v[u], xr, acc[u]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < 8; ++u) sred[warp * 8 + u] = acc[u];
    }
    __syncthreads();
    if (warp == 0 && lane >= c0 && lane < c0 + 8 && lane < w) {
      double tsum = 0.0;
#pragma unroll
      for (int wp = 0; wp < SWEEP_THREADS / 32; ++wp) tsum += sred[wp * 8 + (lane - c0)];
      mine -= tsum;
    }
    __syncthreads();
  }
  if (warp == 0) {
    mine = narrow_bwd_subst(mine, L, R, w, lane);
    if (lane < w) x[I.col0 + lane] = mine;
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(&solved[T.node], 1);
  }
}

__global__ void __launch_bounds__(SWEEP_THREADS) k_fwd_dataflow(
    const SolveCta* __restrict__ ctas, const SolveTask* __restrict__ tasks, const int* __restrict__ targets,
    const int* __restrict__ need, int* __restrict__ done, int* __restrict__ ticket, const SupInfo* __restrict__ sup,
    const int* __restrict__ lR, const double* __restrict__ lv, const double* __restrict__ linv,
    double* __restrict__ y, double* __restrict__ xs, unsigned long long* __restrict__ trace) {
  extern __shared__ __align__(16) double sX[];   // packed lower triangle of the inverse diagonal block (FWD_SWEEP_SMEM)
  __shared__ int s_cta;
  __shared__ double sy[NB_MAX], sx[NB_MAX], spart[SWEEP_THREADS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cta = atomicAdd(ticket, 1);
  __syncthreads();
  const SolveCta C = ctas[s_cta];
  // optional timeline (parsy_cuda_sweep_trace): per CTA the globaltimer at start / inputs ready / end
  if (trace && tid == 0) trace[3 * s_cta] = globaltimer_ns();
  if (C.kind != 1) {
    // narrow supernodes above the leaf region: same task code as k_fwd_narrow, scratch in the dynamic shared memory
    if (C.kind == 2) {
      const SolveTask T = tasks[C.first];
      fwd_narrow_tall_task(T, narrow_info(T), sX, tid, targets, need, done, lR, lv, y, xs);
    } else if (warp < C.count) {
      const SolveTask T = tasks[C.first + warp];
      double* L = sX + warp * NARROW_SCRATCH;
      fwd_narrow_warp_task(T, narrow_info(T), L, L + NARROW_TRI, lane, targets, need, done, lR, lv, y, xs);
    }
    if (trace && tid == 0) { trace[3 * s_cta + 1] = trace[3 * s_cta]; trace[3 * s_cta + 2] = globaltimer_ns(); }
    return;
  }
  const SolveTask T = tasks[C.first];
  const SupInfo I = sup[T.sup];
  const int nb = T.nb, r = I.r, cbase = I.col0 + T.j0;
  // slice of L21 in register tiles of 64 rows: row ri, columns cg*32 .. cg*32+31; the first tile is loaded ahead of
  // the wait, the following ones (same diagonal-block solve, up to SOLVE_TASK_ROWS rows per task) stream behind it
  const int ri = tid & 63, cg = tid >> 6;
  const double* __restrict__ Pbase = lv + I.valptr + (int64_t)T.j0 * r + T.row0;
  const int* __restrict__ rowsT = lR + I.rowptr + T.row0;
  double lval[32];
  int myrow = -1;
  {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const int c = cg * 32 + u;
      lval[u] = (ri < T.nrows && c < nb) ? Pbase[(int64_t)c * r + ri] : 0.0;
    }
    if (cg == 0 && ri < T.nrows) myrow = rowsT[ri];
  }
  // inverse diagonal block: lower triangle packed by columns (column k starts at k*nb - k(k-1)/2), staged with
  // cp.async ahead of the wait so that the post-wait GEMV reads shared memory, not L2
  {
    const double* __restrict__ X = linv + (int64_t)T.slot * NB_MAX * NB_MAX;
    const int i2 = tid & 127;
    for (int k = tid >> 7; k < nb; k += 2)
      if (i2 >= k && i2 < nb) cp_async8(&sX[k * nb - (k * (k - 1)) / 2 + (i2 - k)], X + k * NB_MAX + i2, true);
    cp_async_commit();
  }
  if (tid == 0) spin_until_ge_busy(&done[T.node], need[T.node]);
  if (trace && tid == 0) trace[3 * s_cta + 1] = globaltimer_ns();
  __syncthreads();
  if (tid < nb) sy[tid] = __ldcg(&y[cbase + tid]);
  cp_async_wait<0>();
  __syncthreads();
  // x_b = inv(L_bb) y_b, recomputed by every task of the block column (an extra flag hop would sit on the critical
  // path): row i = tid & 127, the two halves of the block split k
  const int i = tid & 127, half = tid >> 7;
  {
    double acc0 = 0.0, acc1 = 0.0;
    if (i < nb) {
      const int kend = min(i + 1, half ? nb : 64);
      int k = half * 64;
#pragma unroll 4
      for (; k + 1 < kend; k += 2) {
        acc0 = fma(sX[k * nb - (k * (k - 1)) / 2 + (i - k)], sy[k], acc0);
        acc1 = fma(sX[(k + 1) * nb - ((k + 1) * k) / 2 + (i - k - 1)], sy[k + 1], acc1);
      }
      if (k < kend) acc0 = fma(sX[k * nb - (k * (k - 1)) / 2 + (i - k)], sy[k], acc0);
    }
    spart[tid] = acc0 + acc1;
  }
  __syncthreads();
  if (tid < NB_MAX) {   // entries past nb are multiplied by zero-filled lval: they must be zeros, not stale smem
    const double v = (tid < nb) ? spart[tid] + spart[tid + 128] : 0.0;
    sx[tid] = v;
    if (T.first && tid < nb) xs[cbase + tid] = v;
  }
  __syncthreads();
  if (T.nrows > 0) {
    for (int sub = 0; sub < T.nrows; sub += SOLVE_TILE_ROWS) {
      if (sub > 0) {
        const int rr = sub + ri;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const int c = cg * 32 + u;
          lval[u] = (rr < T.nrows && c < nb) ? Pbase[(int64_t)c * r + rr] : 0.0;
        }
        myrow = (cg == 0 && rr < T.nrows) ? rowsT[rr] : -1;
        __syncthreads();   // spart of the previous tile has been consumed
      }
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int u = 0; u < 32; u += 2) {
        t0 = fma(lval[u], sx[(cg * 32 + u) & (NB_MAX - 1)], t0);
        t1 = fma(lval[u + 1], sx[(cg * 32 + u + 1) & (NB_MAX - 1)], t1);
      }
      spart[tid] = t0 + t1;
      __syncthreads();
      if (tid < 64 && myrow >= 0) atomicAdd(&y[myrow], -(spart[tid] + spart[tid + 64] + spart[tid + 128] + spart[tid + 192]));
      // publish this tile: the nodes its rows belong to may be the next link of the chain and must not wait for
      // the rest of the task
      __threadfence();
      __syncthreads();
      const int tile = sub / SOLVE_TILE_ROWS;
      const int* __restrict__ tt = tasks[C.first].tile_tgt;   // from global: no dynamically indexed local copy
      const int qb = tile ? tt[tile - 1] : T.tgt_begin, qe = tt[tile];
      for (int q = qb + tid; q < qe; q += SWEEP_THREADS) atomicAdd(&done[targets[q]], 1);
    }
  }
  if (trace && tid == 0) trace[3 * s_cta + 2] = globaltimer_ns();
}

__global__ void __launch_bounds__(SWEEP_THREADS, 2) k_bwd_dataflow(
    const SolveCta* __restrict__ ctas, int nctas, const SolveTask* __restrict__ tasks,
    const int* __restrict__ targets, const int* __restrict__ ntiles, int* __restrict__ cnt, int* __restrict__ solved,
    int* __restrict__ ticket, const SupInfo* __restrict__ sup, const int* __restrict__ lR,
    const double* __restrict__ lv, const double* __restrict__ linv, double* __restrict__ x,
    unsigned long long* __restrict__ trace) {
  extern __shared__ __align__(16) double sX[];   // packed lower triangle of the inverse diagonal block (FWD_SWEEP_SMEM)
  __shared__ int s_cta, s_last;
  __shared__ double sx[SOLVE_TILE_ROWS], sy[NB_MAX], spart[SWEEP_THREADS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cta = nctas - 1 - atomicAdd(ticket, 1);
  __syncthreads();
  const SolveCta C = ctas[s_cta];
  if (trace && tid == 0) trace[3 * s_cta] = globaltimer_ns();
  if (C.kind != 1) {
    if (C.kind == 2) {
      const SolveTask T = tasks[C.first];
      bwd_narrow_tall_task(T, narrow_info(T), sX, tid, targets, solved, lR, lv, x);
    } else if (warp < C.count) {
      const SolveTask T = tasks[C.first + warp];
      double* L = sX + warp * NARROW_SCRATCH;
      bwd_narrow_warp_task(T, narrow_info(T), L, L + NARROW_TRI, lane, targets, solved, lR, lv, x);
    }
    if (trace && tid == 0) { trace[3 * s_cta + 1] = trace[3 * s_cta]; trace[3 * s_cta + 2] = globaltimer_ns(); }
    return;
  }
  const SolveTask T = tasks[C.first];
  const SupInfo I = sup[T.sup];
  const int nb = T.nb, r = I.r, cbase = I.col0 + T.j0;
  // slice of L21 in register tiles of 64 rows: warp w owns columns w, w+8, ...; lanes own rows lane, lane+32.
  // The first tile is loaded ahead of the wait; partial sums of all tiles of the task accumulate in registers.
  const double* __restrict__ Pbase = lv + I.valptr + (int64_t)T.j0 * r + T.row0;
  const int* __restrict__ rowsT = lR + I.rowptr + T.row0;
  double lval[16][2];
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int c = warp + 8 * u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ii = lane + 32 * h;
      lval[u][h] = (c < nb && ii < T.nrows) ? Pbase[(int64_t)c * r + ii] : 0.0;
    }
  }
  int myrow = -1;
  if (tid < SOLVE_TILE_ROWS && tid < T.nrows) myrow = rowsT[tid];
  // whichever task finishes last solves the diagonal block: stage its inverse (packed lower triangle) ahead of the wait
  {
    const double* __restrict__ X = linv + (int64_t)T.slot * NB_MAX * NB_MAX;
    const int i2 = tid & 127;
    for (int k = tid >> 7; k < nb; k += 2)
      if (i2 >= k && i2 < nb) cp_async8(&sX[k * nb - (k * (k - 1)) / 2 + (i2 - k)], X + k * NB_MAX + i2, true);
    cp_async_commit();
  }
  for (int q = T.tgt_begin + tid; q < T.tgt_end; q += SWEEP_THREADS) spin_until_ge_busy(&solved[targets[q]], 1);
  __syncthreads();
  if (trace && tid == 0) trace[3 * s_cta + 1] = globaltimer_ns();
  if (T.nrows > 0) {
    double part[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) part[u] = 0.0;
    for (int sub = 0; sub < T.nrows; sub += SOLVE_TILE_ROWS) {
      if (sub > 0) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int c = warp + 8 * u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ii = sub + lane + 32 * h;
            lval[u][h] = (c < nb && ii < T.nrows) ? Pbase[(int64_t)c * r + ii] : 0.0;
          }
        }
        myrow = (tid < SOLVE_TILE_ROWS && sub + tid < T.nrows) ? rowsT[sub + tid] : -1;
        __syncthreads();   // sx of the previous tile has been consumed
      }
      if (tid < SOLVE_TILE_ROWS) sx[tid] = (myrow >= 0) ? __ldcg(&x[myrow]) : 0.0;
      __syncthreads();
      const double x0 = sx[lane], x1 = sx[lane + 32];
#pragma unroll
      for (int u = 0; u < 16; ++u) part[u] = fma(lval[u][0], x0, fma(lval[u][1], x1, part[u]));
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int c = warp + 8 * u;
      double p2 = part[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) p2 += __shfl_xor_sync(0xffffffffu, p2, o);
      if (lane == 0 && c < nb) atomicAdd(&x[cbase + c], -p2);
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&cnt[T.node], 1) == ntiles[T.node] - 1);
  __syncthreads();
  if (!s_last) {
    cp_async_wait<0>();
    if (trace && tid == 0) trace[3 * s_cta + 2] = globaltimer_ns();
    return;
  }
  // every slice has added its partial: x_b = inv(L_bb)' x_b, one warp per column of the inverse
  __threadfence();
  if (tid < nb) sy[tid] = __ldcg(&x[cbase + tid]);
  cp_async_wait<0>();
  __syncthreads();
  // x_c = sum_{k >= c} X(k, c) y_k: one thread per column and half of its k range (the same shape as the forward
  // sweep's product; a warp-per-column reduction costs 16 dependent shuffle chains per warp on the critical path)
  {
    const int c = tid & 127, half = tid >> 7;
    double a0 = 0.0, a1 = 0.0;
    if (c < nb) {
      const double* __restrict__ col = sX + (c * nb - (c * (c - 1)) / 2) - c;   // col[k] = X(k, c), k >= c
      const int mid = (c + nb + 1) >> 1;
      int k = half ? mid : c;
      const int kend = half ? nb : mid;
#pragma unroll 4
      for (; k + 1 < kend; k += 2) {
        a0 = fma(col[k], sy[k], a0);
        a1 = fma(col[k + 1], sy[k + 1], a1);
      }
      if (k < kend) a0 = fma(col[k], sy[k], a0);
    }
    spart[tid] = a0 + a1;
  }
  __syncthreads();
  if (tid < nb) x[cbase + tid] = spart[tid] + spart[tid + 128];
  __threadfence();
  __syncthreads();
  if (tid == 0) atomicExch(&solved[T.node], 1);
  if (trace && tid == 0) trace[3 * s_cta + 2] = globaltimer_ns();
}

// ------------------------------------------------------------------------------------------------
// Narrow-only sweep kernels for the leaf region of the tree (the CTAs of the plan that come before the first
// block-column task; on the 2-D problems > 95 % of all supernodes).  Same ticket / counter protocol and the same task
// code as the general kernels, but nothing of the block path: no 66 KB inverse block, ~48 registers, so 40 warps per
// SM are in flight instead of 16 — these tasks are chains of dependent L2/HBM round trips and their throughput is
// occupancy.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SWEEP_THREADS, 5) k_fwd_narrow(
    const SolveCta* __restrict__ ctas, const SolveTask* __restrict__ tasks, const int* __restrict__ targets,
    const int* __restrict__ need, int* __restrict__ done, int* __restrict__ ticket, const SupInfo* __restrict__ sup,
    const int* __restrict__ lR, const double* __restrict__ lv, double* __restrict__ y, double* __restrict__ xs) {
  __shared__ double sScr[8 * NARROW_SCRATCH];
  __shared__ int s_cta;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cta = atomicAdd(ticket, 1);
  __syncthreads();
  const SolveCta C = ctas[s_cta];
  if (C.kind == 2) {
    const SolveTask T = tasks[C.first];
    fwd_narrow_tall_task(T, narrow_info(T), sScr, tid, targets, need, done, lR, lv, y, xs);
    return;
  }
  if (warp >= C.count) return;
  const SolveTask T = tasks[C.first + warp];
  double* L = sScr + warp * NARROW_SCRATCH;
  fwd_narrow_warp_task(T, narrow_info(T), L, L + NARROW_TRI, lane, targets, need, done, lR, lv, y, xs);
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(SWEEP_THREADS, MIN_CTAS) k_bwd_narrow(
    const SolveCta* __restrict__ ctas, int nctas, const SolveTask* __restrict__ tasks, const int* __restrict__ targets,
    int* __restrict__ solved, int* __restrict__ ticket, const SupInfo* __restrict__ sup, const int* __restrict__ lR,
    const double* __restrict__ lv, double* __restrict__ x) {
  __shared__ double sScr[8 * NARROW_SCRATCH];
  __shared__ int s_cta;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cta = nctas - 1 - atomicAdd(ticket, 1);
  __syncthreads();
  const SolveCta C = ctas[s_cta];
  if (C.kind == 2) {
    const SolveTask T = tasks[C.first];
    bwd_narrow_tall_task(T, narrow_info(T), sScr, tid, targets, solved, lR, lv, x);
    return;
  }
  if (warp >= C.count) return;
  const SolveTask T = tasks[C.first + warp];
  double* L = sScr + warp * NARROW_SCRATCH;
  bwd_narrow_warp_task(T, narrow_info(T), L, L + NARROW_TRI, lane, targets, solved, lR, lv, x);
}

// ------------------------------------------------------------------------------------------------
// column (CSC) forward solve (triangularSolve/Triangular_CSC.h:14,50,76), one launch per solve.
//
// One warp per column, diagonal first in every column.  Warps draw tickets and take the columns in the order of the
// caller's schedule (level sets, Triangular_CSC.h:58-70; H-levels x w-partitions, :84-98; or 0..n-1 for the serial
// lsolve) — any topological order of the column DAG.  Column j is ready when indeg[j] updates have arrived
// (indeg = number of off-diagonal entries of row j, counted once per structure); a finished column adds
// -L(i,j) x_j into x_i with red.add, fences, and bumps done[i].  Producers hold smaller tickets than their consumers,
// so spinning cannot deadlock.  Algorithmic traffic: 12 B per stored entry (value + row index) + 16 B per column.
// ------------------------------------------------------------------------------------------------
__global__ void k_csc_indeg(int n, const int* __restrict__ Lp, const int* __restrict__ Li, int* __restrict__ indeg) {
  const int64_t nnz = Lp[n];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    // entry p is the diagonal iff it is the first of its column: binary search for the column
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (Lp[mid] <= p) lo = mid; else hi = mid - 1; }
    if (p != Lp[lo]) atomicAdd(&indeg[Li[p]], 1);
  }
}

__global__ void __launch_bounds__(256) k_csc_dataflow(int n, const int* __restrict__ order, const int* __restrict__ Lp,
                                                       const int* __restrict__ Li, const double* __restrict__ Lx,
                                                       const int* __restrict__ indeg, int* done, int* ticket, double* x) {
  const int lane = threadIdx.x & 31;
  int t = 0;
  if (lane == 0) t = atomicAdd(ticket, 1);
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t >= n) return;
  const int j = order[t];
  const int p0 = Lp[j], p1 = Lp[j + 1];
  // first entries of the column ahead of the wait: the chain then carries one L2 round trip less
  const int pf = p0 + 1 + lane;
  const int i_first = pf < p1 ? Li[pf] : -1;
  const double l_first = pf < p1 ? Lx[pf] : 0.0;
  // lane 0 alone waits, reads x_j and writes the solved value back; the other lanes get x_j by shuffle (they must not
  // read x[j] themselves: lane 0 may already have overwritten it with the solved value)
  double xj = 0.0;
  if (lane == 0) {
    const double dinv = 1.0 / Lx[p0];
    spin_until_ge_busy(&done[j], indeg[j]);
    xj = __ldcg(&x[j]) * dinv;
    x[j] = xj;
  }
  xj = __shfl_sync(0xffffffffu, xj, 0);
  if (i_first >= 0) atomicAdd(&x[i_first], -l_first * xj);
  for (int p = pf + 32; p < p1; p += 32) atomicAdd(&x[Li[p]], -Lx[p] * xj);
  __threadfence();
  __syncwarp();
  if (i_first >= 0) atomicAdd(&done[i_first], 1);
  for (int p = pf + 32; p < p1; p += 32) atomicAdd(&done[Li[p]], 1);
}

// ------------------------------------------------------------------------------------------------
// full system A x = b around the two sweeps: permutation, residual of the symmetric matrix
// given by its lower half, update.  All HBM-bound, algorithmic bytes: 16 B per vector entry moved, 12 B per stored
// entry of A for the residual.
// ------------------------------------------------------------------------------------------------
__global__ void k_perm_gather(int n, const int* __restrict__ perm, const double* __restrict__ b, double* __restrict__ y) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) y[k] = b[perm[k]];
}
__global__ void k_perm_scatter(int n, const int* __restrict__ perm, const double* __restrict__ y, double* __restrict__ x) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) x[perm[k]] = y[k];
}
// res -= A x for the symmetric A whose lower half (by columns) is (c, r, vals); res holds b on entry.
// One thread per column: the column's own contribution is summed in a register, the mirrored ones go out as atomics.
__global__ void k_residual_sym_lower(int n, const int* __restrict__ c, const int* __restrict__ r,
                                     const double* __restrict__ vals, const double* __restrict__ x,
                                     double* __restrict__ res) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double xj = x[j];
    double own = 0.0;
    for (int p = c[j]; p < c[j + 1]; ++p) {
      const int i = r[p];
      const double a = vals[p];
      if (i == j) own = fma(a, xj, own);
      else {
        own = fma(a, x[i], own);            // A(j,i) x(i), the mirrored entry
        atomicAdd(&res[i], -a * xj);        // A(i,j) x(j)
      }
    }
    atomicAdd(&res[j], -own);
  }
}
// out[0] += sum v^2
__global__ void __launch_bounds__(256) k_sumsq(int n, const double* __restrict__ v, double* __restrict__ out) {
  __shared__ double sw[8];
  double acc = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) acc = fma(v[k], v[k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sw[w];
    atomicAdd(out, t);
  }
}
__global__ void k_add_inplace(int n, const double* __restrict__ d, double* __restrict__ x) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) x[k] += d[k];
}

}  // namespace parsy
