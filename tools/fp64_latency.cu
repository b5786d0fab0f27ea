// Dependent-issue latencies (cycles) of the FP64 building blocks that sit on the factorization's critical path
// (pivot chain of k_potrf_block, DMMA accumulate chains, shared-memory / barrier round trips) on one warp of one SM.
// Output: one line per item, cycles per dependent operation.  Used to size the POTRF design.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../parsy_bench_b200/csrc/kernels.cuh"
using namespace parsy;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int N = 512;

// the pivot-chain rsqrt used before rsqrt_pivot(): FP32 seed (MUFU.RSQ) + two Newton steps with range scaling
__device__ __forceinline__ double fast_rsqrt(double a) {
  const int hi = __double2hiint(a);
  const int ex = (((hi >> 20) & 0x7ff) - 1023) & ~1;
  const double as = __hiloint2double(hi - (ex << 20), __double2loint(a));
  double y = (double)rsqrtf((float)as);
  const double h = 0.5 * as;
#pragma unroll
  for (int it = 0; it < 2; ++it) { const double e = fma(-h * y, y, 0.5); y = fma(y, e, y); }
  y = __hiloint2double(__double2hiint(y) - ((ex >> 1) << 20), __double2loint(y));
  if (!(a > 1e-300 && a < 1e300)) y = rsqrt(a);
  return y;
}

#define CHAIN_KERNEL(name, init, body)                                                         \
  __global__ void name(double* out, long long* clk, double seed) {                             \
    double x = seed + threadIdx.x * 1e-12; init;                                               \
    long long t0 = clock64();                                                                  \
    _Pragma("unroll 16") for (int i = 0; i < N; ++i) { body; }                                 \
    long long t1 = clock64();                                                                  \
    out[threadIdx.x] = x; if (threadIdx.x == 0) clk[0] = t1 - t0;                              \
  }

CHAIN_KERNEL(k_dfma, double a = 1.0000001; double b = 1e-9, x = fma(x, a, b))
CHAIN_KERNEL(k_dmul, double a = 1.0000001, x = x * a)
CHAIN_KERNEL(k_dadd, double a = 1e-9, x = x + a)
CHAIN_KERNEL(k_f2f, , x = (double)((float)x) + 0.0)
CHAIN_KERNEL(k_mufu32, , { float f = (float)x; asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f)); x = (double)f; })
CHAIN_KERNEL(k_rsq64h, , { asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x)); })
CHAIN_KERNEL(k_rcp64h, , { asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x)); })
CHAIN_KERNEL(k_fast_rsqrt, , x = fast_rsqrt(x) + 1.0)
CHAIN_KERNEL(k_lib_rsqrt, , x = rsqrt(x) + 1.0)
CHAIN_KERNEL(k_lib_sqrt, , x = sqrt(x) + 1.0)
CHAIN_KERNEL(k_lib_div, double a = 1.5, x = a / x + 1.0)
CHAIN_KERNEL(k_shfl, , x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31))

// the pivot chain's rsqrt (kernels.cuh): MUFU.RSQ64H seed (~2^-22) + one third-order step, no range fix-ups
__device__ __forceinline__ double rsqrt_halley(double a) { return rsqrt_pivot(a); }
CHAIN_KERNEL(k_rsqrt_halley, , x = rsqrt_halley(x) + 1.0)
// two Newton steps from the same seed
__device__ __forceinline__ double rsqrt_newton2(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = 0.5 * a;
#pragma unroll
  for (int it = 0; it < 2; ++it) { const double e = fma(-h * y, y, 0.5); y = fma(y, e, y); }
  return y;
}
CHAIN_KERNEL(k_rsqrt_newton2, , x = rsqrt_newton2(x) + 1.0)

__global__ void k_dmma_acc(double* out, long long* clk) {
  double c0 = 0, c1 = 0, a = 1e-3 * threadIdx.x, b = 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) dmma884(c0, c1, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_dmma_a(double* out, long long* clk) {   // result feeds the next A operand
  double a = 1e-3 * threadIdx.x, b = 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double c0 = 0, c1 = 0; dmma884(c0, c1, a, b); a = c0; }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_smem(double* out, long long* clk) {
  __shared__ double s[64];
  s[threadIdx.x] = threadIdx.x; s[threadIdx.x + 32] = 0;
  __syncwarp();
  double x = 0;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { volatile double* p = s; p[threadIdx.x] = x; __syncwarp(); x = p[(threadIdx.x + 1) & 31] + 1.0; __syncwarp(); }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <int NT> __global__ void k_barrier(double* out, long long* clk) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { __syncthreads(); x += 1.0; }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_gl2(double* buf, double* out, long long* clk) {   // L2 round trip: dependent ld.cg chain
  long long idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 128; ++i) idx = (long long)__ldcg(buf + idx);
  long long t1 = clock64();
  out[threadIdx.x] = (double)idx; if (threadIdx.x == 0) clk[0] = t1 - t0;
}

// 8x8 Cholesky in registers, as every row-thread of potrf_micro8 runs it (all lanes redundant)
template <int VAR> __global__ void k_chol8(double* out, long long* clk, double seed) {
  double d[8][8], rr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int m = 0; m <= k; ++m) d[k][m] = (k == m) ? 12.0 + seed : -1.0 / (1 + k - m) + seed;
  long long t0 = clock64();
  for (int rep = 0; rep < 16; ++rep) {
    if (VAR == 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double piv = d[c][c];
        const double r = fast_rsqrt(piv);
        rr[c] = r; d[c][c] = piv * r;
#pragma unroll
        for (int k = c + 1; k < 8; ++k) d[k][c] *= r;
#pragma unroll
        for (int k = c + 1; k < 8; ++k)
#pragma unroll
          for (int m = c + 1; m <= k; ++m) d[k][m] = fma(-d[k][c], d[m][c], d[k][m]);
      }
    } else if (VAR == 1) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double piv = d[c][c];
        const double r = rsqrt_halley(piv);
        rr[c] = r; d[c][c] = piv * r;
#pragma unroll
        for (int k = c + 1; k < 8; ++k) d[k][c] *= r;
#pragma unroll
        for (int k = c + 1; k < 8; ++k)
#pragma unroll
          for (int m = c + 1; m <= k; ++m) d[k][m] = fma(-d[k][c], d[m][c], d[k][m]);
      }
    } else {
      // 2x2 block pivots: both reciprocal square roots of a column pair start together
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        const double a = d[c][c], b = d[c + 1][c], e = d[c + 1][c + 1];
        const double det = fma(a, e, -b * b);
        const double r0 = rsqrt_halley(a);          // 1/l00
        const double rd = rsqrt_halley(det);        // 1/sqrt(det)
        const double l00 = a * r0;
        const double r1 = rd * l00;                 // 1/l11 = sqrt(a)/sqrt(det)
        rr[c] = r0; rr[c + 1] = r1;
        d[c][c] = l00; d[c + 1][c] = b * r0; d[c + 1][c + 1] = det * rd * r0;   // sqrt(det)/sqrt(a)
        const double l10 = d[c + 1][c];
#pragma unroll
        for (int k = c + 2; k < 8; ++k) {
          const double x0 = d[k][c] * r0;
          const double x1 = fma(-x0, l10, d[k][c + 1]) * r1;
          d[k][c] = x0; d[k][c + 1] = x1;
        }
#pragma unroll
        for (int k = c + 2; k < 8; ++k)
#pragma unroll
          for (int m = c + 2; m <= k; ++m) d[k][m] = fma(-d[k][c + 1], d[m][c + 1], fma(-d[k][c], d[m][c], d[k][m]));
      }
    }
    // feed the result back so that repetitions stay dependent and SPD
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int m = 0; m <= k; ++m) d[k][m] = (k == m) ? 12.0 + 1e-3 * d[k][m] : -1.0 / (1 + k - m) + 1e-3 * d[k][m];
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { s += rr[k];
#pragma unroll
    for (int m = 0; m <= k; ++m) s += d[k][m]; }
  out[threadIdx.x] = s; if (threadIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  double* out; long long* clk; double* buf;
  CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMalloc(&clk, 64)); CK(cudaMalloc(&buf, 4096 * 8));
  { double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (double)((i * 33 + 97) & 4095); CK(cudaMemcpy(buf, h, sizeof(h), cudaMemcpyHostToDevice)); }
  long long c;
#define RUN(label, per, launch) do { launch; launch; CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost)); \
    printf("{\"item\": \"%s\", \"cycles\": %.1f}\n", label, (double)c / (per)); } while (0)
  RUN("dfma_dep", N, (k_dfma<<<1, 32>>>(out, clk, 1.0)));
  RUN("dmul_dep", N, (k_dmul<<<1, 32>>>(out, clk, 1.0)));
  RUN("dadd_dep", N, (k_dadd<<<1, 32>>>(out, clk, 1.0)));
  RUN("f2f_f64_f32_f64_plus_dadd", N, (k_f2f<<<1, 32>>>(out, clk, 1.5)));
  RUN("cvt+mufu.rsq.f32+cvt", N, (k_mufu32<<<1, 32>>>(out, clk, 1.5)));
  RUN("rsqrt.approx.ftz.f64", N, (k_rsq64h<<<1, 32>>>(out, clk, 1.5)));
  RUN("rcp.approx.ftz.f64", N, (k_rcp64h<<<1, 32>>>(out, clk, 1.5)));
  RUN("fast_rsqrt(current)+dadd", N, (k_fast_rsqrt<<<1, 32>>>(out, clk, 1.5)));
  RUN("rsqrt_halley+dadd", N, (k_rsqrt_halley<<<1, 32>>>(out, clk, 1.5)));
  RUN("rsqrt_newton2+dadd", N, (k_rsqrt_newton2<<<1, 32>>>(out, clk, 1.5)));
  RUN("lib_rsqrt+dadd", N, (k_lib_rsqrt<<<1, 32>>>(out, clk, 1.5)));
  RUN("lib_sqrt+dadd", N, (k_lib_sqrt<<<1, 32>>>(out, clk, 1.5)));
  RUN("lib_div+dadd", N, (k_lib_div<<<1, 32>>>(out, clk, 1.5)));
  RUN("shfl_f64", N, (k_shfl<<<1, 32>>>(out, clk, 1.5)));
  RUN("dmma884_acc_dep", N, (k_dmma_acc<<<1, 32>>>(out, clk)));
  RUN("dmma884_a_dep", N, (k_dmma_a<<<1, 32>>>(out, clk)));
  RUN("smem_st_ld_roundtrip(2 syncwarp)+dadd", N, (k_smem<<<1, 32>>>(out, clk)));
  RUN("syncthreads_128", N, (k_barrier<128><<<1, 128>>>(out, clk)));
  RUN("syncthreads_256", N, (k_barrier<256><<<1, 256>>>(out, clk)));
  RUN("syncthreads_512", N, (k_barrier<512><<<1, 512>>>(out, clk)));
  RUN("ldcg_l2_dep", 128, (k_gl2<<<1, 32>>>(buf, out, clk)));
  RUN("chol8_regs_current(per 8x8)", 16, (k_chol8<0><<<1, 32>>>(out, clk, 0.0)));
  RUN("chol8_regs_halley(per 8x8)", 16, (k_chol8<1><<<1, 32>>>(out, clk, 0.0)));
  RUN("chol8_regs_2x2pivot(per 8x8)", 16, (k_chol8<2><<<1, 32>>>(out, clk, 0.0)));
  // accuracy of the candidate rsqrt against the library
  return 0;
}
