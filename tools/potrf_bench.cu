// Phase timing of k_potrf_block's building blocks on one SM (clock64 around each phase).
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
  potrf_in_smem(S, rd, nb, nbp, tid, info, 0);
  long long t2 = clock64();
  store_factor_block(S, P, r, nb, tid);
  long long t3 = clock64();
  invert_in_smem(S, rd, XD, Tt, nbp, tid);
  long long t4 = clock64();
  store_inverse(S, XD, linv, nb, tid);
  __syncthreads();
  long long t5 = clock64();
  if (tid == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; clk[4] = t5 - t4; }
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
    potrf_mid8(S, p0, nbp, tid); __syncthreads();
    long long t2 = clock64();
    potrf_micro8(S, rd, p0 + 8, nbp, tid, info, 0, nb); __syncthreads();
    long long t3 = clock64();
    if (p0 + 16 < nbp) { potrf_trailing16(S, p0, nbp, warp, lane); __syncthreads(); }
    long long t4 = clock64();
    a += t1 - t0; b += t2 - t1; c += t3 - t2; d += t4 - t3;
  }
  if (tid == 0) { clk[0] = a; clk[1] = b; clk[2] = c; clk[3] = d; }
}
int main() {
  const int nb = 128, r = 1000;
  std::vector<double> A((size_t)r * nb, 0.0);
  for (int c = 0; c < nb; ++c) for (int i = c; i < nb; ++i) A[(size_t)c * r + i] = (i == c) ? 12.0 : -1.0 / (1 + (i - c));
  double *dP, *dX; int* dinfo; long long* dclk;
  cudaMalloc(&dP, A.size() * 8); cudaMalloc(&dX, 128 * 128 * 8); cudaMalloc(&dinfo, 4); cudaMalloc(&dclk, 64);
  cudaMemset(dinfo, 0, 4);
  cudaFuncSetAttribute(k_phases, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
  cudaFuncSetAttribute(k_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
  for (int it = 0; it < 3; ++it) {
    cudaMemcpy(dP, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_phases<<<1, POTRF_THREADS, POTRF_SMEM>>>(dP, r, nb, dX, dinfo, dclk);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk[8]; cudaMemcpy(clk, dclk, 40, cudaMemcpyDeviceToHost);
    int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
    printf("total %.1f us  load %lld  chol %lld  writeback %lld  invert %lld  store_inv %lld  (clk) info %d err %s\n", ms * 1e3, clk[0], clk[1], clk[2], clk[3], clk[4], info, cudaGetErrorString(cudaGetLastError()));
  }
  cudaMemcpy(dP, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  k_sub<<<1, POTRF_THREADS, POTRF_SMEM>>>(dP, r, nb, dinfo, dclk);
  cudaDeviceSynchronize();
  long long clk[8]; cudaMemcpy(clk, dclk, 32, cudaMemcpyDeviceToHost);
  printf("chol split: micro8(a) %lld  mid8 %lld  micro8(b) %lld  trailing16 %lld (clk, summed over 8 macro panels)\n", clk[0], clk[1], clk[2], clk[3]);
  return 0;
}
