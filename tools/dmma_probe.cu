// tools/dmma_probe.cu -- microbenchmark (GPU box): FP64 tensor-core (DMMA, mma.sync.aligned.m8n8k4.f64) throughput of the
// device next to the FP64 FMA pipe, the two candidate engines of the shared-KKT ADMM path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dmma(double *out, int iters) {
  // 8 independent accumulator tiles per warp: enough ILP to cover the MMA latency
  double c[8][2];
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i * 1e-9; }
  const double a = 1.0000001, b = 0.9999999;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma_chain(double *out, int iters, long long *cyc) {
  double c0 = threadIdx.x * 1e-9, c1 = 1e-9;
  const double a = 1.0000001, b = 0.9999999;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 13;
  double *out; long long *cyc, h;
  cudaMalloc(&out, (size_t)blocks * threads * 8); cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // one m8n8k4 = 8*8*4 FMAs = 512 flops per warp
    double fl = 512.0 * 8 * (double)iters * blocks * (threads / 32);
    if (rep) printf("DMMA m8n8k4.f64 : %.2f TFLOP/s (%.3f ms)\n", fl / (ms * 1e-3) / 1e12, ms);
    cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters * 4); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 8 * (double)iters * 4 * blocks * threads;
    if (rep) printf("DFMA            : %.2f TFLOP/s (%.3f ms)\n", fl / (ms * 1e-3) / 1e12, ms);
  }
  k_dmma_chain<<<1, 32>>>(out, 4096, cyc); k_dmma_chain<<<1, 32>>>(out, 4096, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA dependent chain: %.1f cycles per mma\n", (double)h / 4096);
  printf("%s, %d SMs, cudaError=%s\n", p.name, p.multiProcessorCount, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
