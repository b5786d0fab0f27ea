// FP64 roofline microbenchmarks for B200 (sm_100a): register-resident DMMA issue rate for each
// mma.sync f64 shape, plain DFMA rate, and red.global.add.f64 throughput.  MEASURED_PEAKS.json has
// no FP64 entry, so the factorization roofline denominator comes from here.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP> __global__ void k_m8n8k4(double* out, int iters) {
  double c[ILP][2]; double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
  for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ILP> __global__ void k_m16n8k4(double* out, int iters) {
  double c[ILP][4]; double a0 = threadIdx.x * 1e-9, a1 = 0.5, b = 1.0 + threadIdx.x * 1e-9;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
  }
  double s = 0; for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ILP> __global__ void k_m16n8k8(double* out, int iters) {
  double c[ILP][4]; double a0 = threadIdx.x * 1e-9, a1 = 0.5, a2 = 0.25, a3 = 0.125, b0 = 1.0 + threadIdx.x * 1e-9, b1 = 0.3;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0; for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ILP> __global__ void k_m16n8k16(double* out, int iters) {
  double c[ILP][4]; double a[8], b[4];
  for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 1e-9 + j; for (int j = 0; j < 4; j++) b[j] = 1.0 / (j + 1 + threadIdx.x);
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0; for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ILP> __global__ void k_dfma(double* out, int iters) {
  double c[ILP]; double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int i = 0; i < ILP; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0; for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// red.add.f64: each thread adds to a distinct-ish address; stride pattern chosen by mode
__global__ void k_red(double* buf, size_t n, int iters, int mode) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; it++) {
    size_t idx;
    if (mode == 0) idx = (tid + (size_t)it * nt) % n;                    // coalesced, distinct
    else { size_t w = tid / 32, l = tid % 32; idx = ((w * 2654435761ull + it * 40503ull) % (n / 4096)) * 4096 + l * 128 % 4096 + (l / 32); } // 32 separate 128B lines per warp
    atomicAdd(&buf[idx], -1.0);
  }
}
__global__ void k_rmw(double* buf, size_t n, int iters) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; it++) { size_t idx = (tid + (size_t)it * nt) % n; buf[idx] -= 1.0; }
}

template<typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount; printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    int threads = warps * 32; int grid = sms * (warps >= 16 ? 1 : 2);
    double nw = (double)grid * warps;
#define RUN(name, kern, ILP, flops_per_mma) { float ms = timeit([&]{ kern<ILP><<<grid, threads>>>(out, iters); }); \
      printf("{\"bench\": \"%s\", \"ilp\": %d, \"warps_per_cta\": %d, \"ctas\": %d, \"tflops\": %.2f}\n", name, ILP, warps, grid, nw * iters * ILP * (double)(flops_per_mma) / (ms * 1e-3) / 1e12); }
    RUN("dmma_m8n8k4", k_m8n8k4, 8, 2 * 8 * 8 * 4);
    RUN("dmma_m8n8k4", k_m8n8k4, 16, 2 * 8 * 8 * 4);
    RUN("dmma_m16n8k4", k_m16n8k4, 8, 2 * 16 * 8 * 4);
    RUN("dmma_m16n8k8", k_m16n8k8, 8, 2 * 16 * 8 * 8);
    RUN("dmma_m16n8k16", k_m16n8k16, 8, 2 * 16 * 8 * 16);
    RUN("dfma", k_dfma, 16, 2 * 32);
  }
  size_t n = (size_t)1 << 28; double* buf; CK(cudaMalloc(&buf, n * 8)); CK(cudaMemset(buf, 0, n * 8));
  for (int mode = 0; mode < 2; mode++) {
    int grid = sms * 8, threads = 256, it = 256;
    float ms = timeit([&]{ k_red<<<grid, threads>>>(buf, n, it, mode); });
    printf("{\"bench\": \"red_f64\", \"mode\": %d, \"gatomics_per_s\": %.2f}\n", mode, (double)grid * threads * it / (ms * 1e-3) / 1e9);
  }
  { int grid = sms * 8, threads = 256, it = 256; float ms = timeit([&]{ k_rmw<<<grid, threads>>>(buf, n, it); });
    printf("{\"bench\": \"rmw_f64\", \"gops_per_s\": %.2f}\n", (double)grid * threads * it / (ms * 1e-3) / 1e9); }
  return 0;
}
