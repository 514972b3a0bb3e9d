// spectral_b200/csrc/corridor.cu -- __global__ wrapper of the corridor kernel (K1 + K2).
// Built with --fmad=false: the reference's corridor arithmetic is SSE2 without FMA contraction and
// the segments must be bit-exact (see corridor.cuh).
#include <cuda_runtime.h>

#include <mutex>

#include "bounds.cuh"
#include "corridor.cuh"

__global__ void k_corridor(const CorridorArgs a) {
  extern __shared__ __align__(16) unsigned char corridor_smem[];
  const int b = blockIdx.x;
  corridor_cta_body(a, b, threadIdx.x >> 5, threadIdx.x & 31, corridor_smem, []() { __syncthreads(); });
}

// cudaFuncSetAttribute applies to the CURRENT device and to the FUNCTION, not to a handle: two handles on one device
// share the opt-in, so a handle that needs less must never lower what another one raised (that made the next R = 8
// launch of the other handle fail with "invalid argument").  The size configured per device is kept here, under a
// mutex (handles may live on different host threads); *configured mirrors it for the handle's diagnostics.
static std::mutex g_corridor_mu;
static int g_corridor_configured[64] = {};
extern "C" int spectral_corridor_prepare(int N, int R, int *configured) {
  const CorridorSmem L = corridor_smem_layout(N, R);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  std::lock_guard<std::mutex> lock(g_corridor_mu);
  if (L.total > g_corridor_configured[dev]) {
    if (cudaFuncSetAttribute(k_corridor, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total) != cudaSuccess) return -1;
    g_corridor_configured[dev] = L.total;
  }
  *configured = g_corridor_configured[dev];
  return 0;
}

extern "C" void spectral_launch_corridor(const CorridorArgs &a, cudaStream_t st) {
  const CorridorSmem L = corridor_smem_layout(a.N, a.R);
  k_corridor<<<a.B, 32 * a.R, L.total, st>>>(a);
}

// upstream bound generator (bounds.cuh): one warp per scenario, four warps per CTA, grid-stride over scenarios.  Per warp
// a lane table + a [4][N] line cache in dynamic shared memory (3.6 KB at N = 71), registers capped for 8 CTAs per SM: the
// kernel is latency-bound (sequential table logic, IEEE divisions), so resident warps are what it needs.
__global__ void __launch_bounds__(128, 8) k_bounds(const BoundsArgs a) {
  extern __shared__ __align__(16) unsigned char bounds_smem[];
  const int warp = threadIdx.x >> 5;
  unsigned char *mine = bounds_smem + (size_t)warp * ((spb_smem_bytes_per_warp(a.N) + 15) / 16 * 16);
  SpbTable *T = reinterpret_cast<SpbTable *>(mine);
  double *line = reinterpret_cast<double *>(mine + sizeof(SpbTable));
  for (int b = blockIdx.x * 4 + warp; b < a.B; b += gridDim.x * 4) bounds_warp_body(a, b, threadIdx.x & 31, T, line);
}
extern "C" int spectral_launch_bounds(const BoundsArgs &a, int sm_count, cudaStream_t st) {
  const size_t smem = 4 * ((spb_smem_bytes_per_warp(a.N) + 15) / 16 * 16);
  if (smem > 48 * 1024 && cudaFuncSetAttribute(k_bounds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  int blocks = (a.B + 3) / 4;
  if (blocks > 8 * sm_count) blocks = 8 * sm_count;
  k_bounds<<<blocks, 128, smem, st>>>(a);
  return 0;
}
