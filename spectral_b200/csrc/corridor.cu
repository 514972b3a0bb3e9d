// spectral_b200/csrc/corridor.cu -- __global__ wrapper of the corridor kernel (K1 + K2).
// Built with --fmad=false: the reference's corridor arithmetic is SSE2 without FMA contraction and
// the segments must be bit-exact (see corridor.cuh).
#include <cuda_runtime.h>

#include "corridor.cuh"

__global__ void k_corridor(const CorridorArgs a) {
  extern __shared__ __align__(16) unsigned char corridor_smem[];
  const int b = blockIdx.x;
  corridor_cta_body(a, b, threadIdx.x >> 5, threadIdx.x & 31, corridor_smem, []() { __syncthreads(); });
}

// cudaFuncSetAttribute applies to the CURRENT device: the caller (one handle per GPU) has set its device and remembers
// the size it configured in *configured (per handle, so several handles / devices / threads in one process are fine).
extern "C" int spectral_corridor_prepare(int N, int R, int *configured) {
  const CorridorSmem L = corridor_smem_layout(N, R);
  if (L.total > *configured) {
    if (cudaFuncSetAttribute(k_corridor, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total) != cudaSuccess) return -1;
    *configured = L.total;
  }
  return 0;
}

extern "C" void spectral_launch_corridor(const CorridorArgs &a, cudaStream_t st) {
  const CorridorSmem L = corridor_smem_layout(a.N, a.R);
  k_corridor<<<a.B, 32 * a.R, L.total, st>>>(a);
}
