// k_potrf_block on one SM: correctness against a host Cholesky / triangular inverse for several block widths, and
// clock64 phase timing of its building blocks for one 128x128 SPD block (one CTA, 256 threads).
#include <cstdio>
#include <vector>
#include <cmath>
#include "../parsy_bench_b200/csrc/kernels.cuh"
using namespace parsy;
__global__ void __launch_bounds__(POTRF_THREADS) k_phases(double* P, int r, int nb, double* linv, int* info, long long* clk) {
  extern __shared__ __align__(16) double smem[];
  double* S = smem; double* rd = S + POTRF_S; double* XD = rd + NB_MAX; double* Tt = XD + POTRF_XD;
  const int tid = threadIdx.x, nbp = (nb + 15) & ~15;
  long long t0 = clock64();
  load_padded_block(S, P, r, nb, nbp, tid);
  __syncthreads();
  long long t1 = clock64();
  potrf_in_smem(S, rd, nb, nbp, tid, info, 0, P, r);
  long long t2 = clock64();
  invert_in_smem(S, rd, XD, Tt, linv, nb, nbp, tid);
  __syncthreads();
  long long t3 = clock64();
  if (tid == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; }
}
__global__ void __launch_bounds__(POTRF_THREADS) k_sub(double* P, int r, int nb, int* info, long long* clk) {
  extern __shared__ __align__(16) double smem[];
  double* S = smem; double* rd = S + POTRF_S;
  const int tid = threadIdx.x, nbp = (nb + 15) & ~15, warp = tid >> 5, lane = tid & 31;
  load_padded_block(S, P, r, nb, nbp, tid);
  __syncthreads();
  long long a = 0, b = 0, c = 0, d = 0;
  for (int p0 = 0; p0 < nbp; p0 += 16) {
    long long t0 = clock64();
    potrf_micro8(S, rd, p0, nbp, tid, info, 0, nb); __syncthreads();
    long long t1 = clock64();
    potrf_mid8(S, p0, nbp, warp, lane); __syncthreads();
    long long t2 = clock64();
    potrf_micro8(S, rd, p0 + 8, nbp, tid, info, 0, nb); __syncthreads();
    long long t3 = clock64();
    if (p0 + 16 < nbp) { potrf_trailing16(S, p0, nbp, warp, lane); __syncthreads(); }
    long long t4 = clock64();
    a += t1 - t0; b += t2 - t1; c += t3 - t2; d += t4 - t3;
  }
  if (tid == 0) { clk[0] = a; clk[1] = b; clk[2] = c; clk[3] = d; }
}

static void host_chol_inv(int nb, int r, const std::vector<double>& A, std::vector<long double>& L, std::vector<long double>& X) {
  L.assign((size_t)nb * nb, 0.0L); X.assign((size_t)nb * nb, 0.0L);
  for (int c = 0; c < nb; ++c) {
    for (int i = c; i < nb; ++i) {
      long double v = A[(size_t)c * r + i];
      for (int k = 0; k < c; ++k) v -= L[(size_t)k * nb + i] * L[(size_t)k * nb + c];
      L[(size_t)c * nb + i] = (i == c) ? sqrtl(v) : v / L[(size_t)c * nb + c];
    }
  }
  for (int c = 0; c < nb; ++c)
    for (int i = c; i < nb; ++i) {
      long double v = (i == c) ? 1.0L : 0.0L;
      for (int k = c; k < i; ++k) v -= L[(size_t)k * nb + i] * X[(size_t)c * nb + k];
      X[(size_t)c * nb + i] = v / L[(size_t)i * nb + i];
    }
}

int main() {
  int* dinfo; long long* dclk; double* dX;
  cudaMalloc(&dinfo, 4); cudaMalloc(&dclk, 64); cudaMalloc(&dX, 128 * 128 * 8);
  cudaFuncSetAttribute(k_potrf_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
  cudaFuncSetAttribute(k_phases, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
  cudaFuncSetAttribute(k_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
  // ---- correctness of the production kernel --------------------------------------------------------------------
  const int widths[] = {128, 127, 100, 64, 37, 16, 9, 1};
  bool all_ok = true;
  for (int nb : widths) {
    const int r = nb + 40, j0 = 0;
    std::vector<double> A((size_t)r * nb, 0.0);
    for (int c = 0; c < nb; ++c)
      for (int i = c; i < r; ++i) A[(size_t)c * r + i] = (i == c) ? 6.0 + 0.01 * c : -0.04 / (1 + ((i - c) % 7)) * ((i * 7 + c * 3) % 5 == 0 ? 0.0 : 1.0);
    std::vector<long double> L, X;
    host_chol_inv(nb, r, A, L, X);
    double* dP; cudaMalloc(&dP, A.size() * 8);
    cudaMemcpy(dP, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dX, 0, 128 * 128 * 8); cudaMemset(dinfo, 0, 4);
    SupInfo si; si.rowptr = 0; si.valptr = 0; si.col0 = 0; si.w = nb; si.r = r; si.flags = 0;
    BlockTask bt; bt.sup = 0; bt.j0 = j0; bt.nb = nb; bt.slot = 0; bt.tile0 = 0;
    SupInfo* dsi; BlockTask* dbt; cudaMalloc(&dsi, sizeof(si)); cudaMalloc(&dbt, sizeof(bt));
    cudaMemcpy(dsi, &si, sizeof(si), cudaMemcpyHostToDevice); cudaMemcpy(dbt, &bt, sizeof(bt), cudaMemcpyHostToDevice);
    const int cols = (nb + 15) & ~15;
    k_potrf_block<<<1, POTRF_THREADS, potrf_smem_bytes(cols)>>>(dbt, dsi, dP, dX, dinfo, cols);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> out(A.size()), xo(128 * 128);
    cudaMemcpy(out.data(), dP, A.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(xo.data(), dX, 128 * 128 * 8, cudaMemcpyDeviceToHost);
    int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
    double el = 0, ex = 0, eu = 0, erest = 0;
    for (int c = 0; c < nb; ++c)
      for (int i = 0; i < r; ++i) {
        const double got = out[(size_t)c * r + i];
        if (i < c) eu = fmax(eu, fabs(got - A[(size_t)c * r + i]));                     // strictly upper: untouched
        else if (i < nb) el = fmax(el, fabs(got - (double)L[(size_t)c * nb + i]) / fmax(1e-3, fabs((double)L[(size_t)c * nb + i])));
        else erest = fmax(erest, fabs(got - A[(size_t)c * r + i]));                     // rows below the block: untouched
      }
    for (int c = 0; c < 128; ++c)
      for (int i = 0; i < 128; ++i) {
        const double want = (c < nb && i < nb && i >= c) ? (double)X[(size_t)c * nb + i] : 0.0;
        ex = fmax(ex, fabs(xo[(size_t)c * 128 + i] - want) / fmax(1e-3, fabs(want)));
      }
    const bool ok = e == cudaSuccess && info == 0 && el < 1e-13 && ex < 1e-12 && eu == 0.0 && erest == 0.0;
    all_ok = all_ok && ok;
    printf("nb %3d: err %s info %d  max rel err L %.2e  inv %.2e  upper touched %.1e  below touched %.1e  %s\n", nb,
           cudaGetErrorString(e), info, el, ex, eu, erest, ok ? "OK" : "FAIL");
    cudaFree(dP); cudaFree(dsi); cudaFree(dbt);
  }
  // a non-SPD block must report its column
  {
    const int nb = 64, r = 64;
    std::vector<double> A((size_t)r * nb, 0.0);
    for (int c = 0; c < nb; ++c) A[(size_t)c * r + c] = (c == 37) ? -1.0 : 4.0;
    double* dP; cudaMalloc(&dP, A.size() * 8); cudaMemcpy(dP, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dinfo, 0, 4);
    SupInfo si; si.rowptr = 0; si.valptr = 0; si.col0 = 100; si.w = nb; si.r = r; si.flags = 0;
    BlockTask bt; bt.sup = 0; bt.j0 = 0; bt.nb = nb; bt.slot = 0; bt.tile0 = 0;
    SupInfo* dsi; BlockTask* dbt; cudaMalloc(&dsi, sizeof(si)); cudaMalloc(&dbt, sizeof(bt));
    cudaMemcpy(dsi, &si, sizeof(si), cudaMemcpyHostToDevice); cudaMemcpy(dbt, &bt, sizeof(bt), cudaMemcpyHostToDevice);
    k_potrf_block<<<1, POTRF_THREADS, potrf_smem_bytes(64)>>>(dbt, dsi, dP, dX, dinfo, 64);
    cudaDeviceSynchronize();
    int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
    printf("non-SPD block: info %d (expect %d) %s\n", info, 100 + 37 + 1, info == 138 ? "OK" : "FAIL");
    all_ok = all_ok && info == 138;
  }
  // ---- phase timing ----------------------------------------------------------------------------------------------
  const int nb = 128, r = 1000;
  std::vector<double> A((size_t)r * nb, 0.0);
  for (int c = 0; c < nb; ++c) for (int i = c; i < nb; ++i) A[(size_t)c * r + i] = (i == c) ? 12.0 : -1.0 / (1 + (i - c));
  double* dP; cudaMalloc(&dP, A.size() * 8);
  for (int it = 0; it < 3; ++it) {
    cudaMemcpy(dP, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dinfo, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_phases<<<1, POTRF_THREADS, POTRF_SMEM>>>(dP, r, nb, dX, dinfo, dclk);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk[8]; cudaMemcpy(clk, dclk, 24, cudaMemcpyDeviceToHost);
    int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
    printf("total %.1f us  load %lld  chol(+panel stores) %lld  invert(+stores) %lld  (cycles) info %d err %s\n", ms * 1e3, clk[0], clk[1], clk[2], info, cudaGetErrorString(cudaGetLastError()));
  }
  cudaMemcpy(dP, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  k_sub<<<1, POTRF_THREADS, POTRF_SMEM>>>(dP, r, nb, dinfo, dclk);
  cudaDeviceSynchronize();
  long long clk[8]; cudaMemcpy(clk, dclk, 32, cudaMemcpyDeviceToHost);
  printf("chol split (summed over the 8 macro panels): micro8(a) %lld  mid8 %lld  micro8(b) %lld  trailing16 %lld (cycles)\n", clk[0], clk[1], clk[2], clk[3]);
  printf("%s\n", all_ok ? "ALL OK" : "SOME FAILED");
  return all_ok ? 0 : 1;
}
