// tools/lat_probe.cu -- microbenchmark (GPU box): dependent-issue latency of the FP64 pipe, shared-memory loads,
// shuffles and CTA barriers on this device.  nvcc -arch=sm_100a -O3 -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_chain(double *out, int n, long long *cyc) {
  double a = threadIdx.x * 1e-9 + 1.0, m = 1.0000001, c = 1e-7;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = fma(a, m, c); a = fma(a, m, c); a = fma(a, m, c); a = fma(a, m, c); }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dadd_chain(double *out, int n, long long *cyc) {
  double a = threadIdx.x * 1e-9 + 1.0, c = 1e-7;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = a + c; a = a - c * 0.5; a = a + c; a = a - c * 0.25; }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma_ilp4(double *out, int n, long long *cyc) {
  double a = threadIdx.x * 1e-9 + 1.0, b = a + 1, d = a + 2, e = a + 3, m = 1.0000001, c = 1e-7;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = fma(a, m, c); b = fma(b, m, c); d = fma(d, m, c); e = fma(e, m, c); }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = a + b + d + e;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds_chain(double *out, int n, long long *cyc) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 33 + 1) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { p = idx[p]; p = idx[p]; p = idx[p]; p = idx[p]; }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = p;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl_chain(double *out, int n, long long *cyc) {
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = __shfl_xor_sync(0xffffffffu, a, 1); a = __shfl_xor_sync(0xffffffffu, a, 2); a = __shfl_xor_sync(0xffffffffu, a, 1); a = __shfl_xor_sync(0xffffffffu, a, 2); }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_bar(double *out, int n, long long *cyc) {
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = 1.0;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dsetp_sel(double *out, int n, long long *cyc) {
  double a = threadIdx.x * 1e-3, lo = 0.25, hi = 0.75;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = fmin(fmax(a * 1.0001, lo), hi); a = fmin(fmax(a * 1.0001, lo), hi); }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int n = 4096;
  struct { const char *name; void (*k)(double *, int, long long *); int ops; } tests[] = {
      {"DFMA dependent chain", k_dfma_chain, 4}, {"DADD dependent chain", k_dadd_chain, 4}, {"DFMA 4 independent chains", k_dfma_ilp4, 4},
      {"LDS.32 dependent chain", k_lds_chain, 4}, {"SHFL.64 dependent chain", k_shfl_chain, 4}, {"BAR.SYNC", k_bar, 4},
      {"DMUL+clip(fmax,fmin) chain", k_dsetp_sel, 2}};
  for (auto &t : tests) {
    for (int threads : {32, 128, 192, 512}) {
      t.k<<<1, threads>>>(out, n, cyc);  // warm
      t.k<<<1, threads>>>(out, n, cyc);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-32s threads=%4d : %.1f cycles per op\n", t.name, threads, (double)h / (n * t.ops));
    }
  }
  return 0;
}
