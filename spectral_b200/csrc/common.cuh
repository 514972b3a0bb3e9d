// spectral_b200/csrc/common.cuh -- shared device-side definitions for the sm_100a kernels.
//
// The kernels are written as "warp bodies": __device__ functions that take (warp id, lane id, a
// shared-memory slab) so that tests/warp_emu can compile the very same source with g++ and run one
// warp as 32 lock-stepped host threads (debugging aid for the kernel LOGIC only; the product
// library contains no CPU path and fails loudly without a CUDA device).
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/spectral.h"

#ifdef SPECTRAL_CPU_EMU
#include "warp_emu.h"  // provides SP_DEV, sp_* warp primitives, rn_* arithmetic
#else
#include <cuda_runtime.h>
#define SP_DEV __device__ __forceinline__
#define SP_DEV_NOINLINE __device__ __noinline__
#define SP_HD __host__ __device__ inline
#define SP_FULL 0xffffffffu

SP_DEV double sp_shfl(double v, int src) { return __shfl_sync(SP_FULL, v, src); }
SP_DEV int sp_shfl_i(int v, int src) { return __shfl_sync(SP_FULL, v, src); }
SP_DEV double sp_shfl_up(double v, int d, int width) { return __shfl_up_sync(SP_FULL, v, d, width); }
SP_DEV double sp_shfl_down(double v, int d, int width) { return __shfl_down_sync(SP_FULL, v, d, width); }
SP_DEV double sp_shfl_xor(double v, int m) { return __shfl_xor_sync(SP_FULL, v, m); }
SP_DEV int sp_shfl_xor_i(int v, int m) { return __shfl_xor_sync(SP_FULL, v, m); }
SP_DEV int sp_shfl_up_i(int v, int d, int width) { return __shfl_up_sync(SP_FULL, v, d, width); }
SP_DEV unsigned sp_ballot(int pred) { return __ballot_sync(SP_FULL, pred); }
SP_DEV int sp_any(int pred) { return __any_sync(SP_FULL, pred); }
SP_DEV int sp_all(int pred) { return __all_sync(SP_FULL, pred); }
SP_DEV void sp_syncwarp() { __syncwarp(); }
SP_DEV void sp_sync_cta() { __syncthreads(); }
SP_DEV int sp_warp_in_cta() { return (int)(threadIdx.x >> 5); }
SP_DEV int sp_popc(unsigned v) { return __popc(v); }
SP_DEV int sp_ffs(unsigned v) { return __ffs(v); }
// IEEE round-to-nearest without FMA contraction: the reference is x86-64 SSE2 code compiled
// without FMA, and its corridor / bound arithmetic must be reproduced bit for bit.
SP_DEV double rn_add(double a, double b) { return __dadd_rn(a, b); }
SP_DEV double rn_sub(double a, double b) { return __dsub_rn(a, b); }
SP_DEV double rn_mul(double a, double b) { return __dmul_rn(a, b); }
SP_DEV double rn_div(double a, double b) { return __ddiv_rn(a, b); }
SP_DEV void sp_store2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }  // p 16-byte aligned
#endif

// group (sub-warp) reductions over `width` consecutive lanes, width a power of two
SP_DEV double sp_group_max(double v, int width) {
  for (int m = width >> 1; m > 0; m >>= 1) v = fmax(v, sp_shfl_xor(v, m));
  return v;
}
SP_DEV double sp_group_sum(double v, int width) {
  for (int m = width >> 1; m > 0; m >>= 1) v = v + sp_shfl_xor(v, m);
  return v;
}
// warp-wide maximum of NON-NEGATIVE doubles (norms): their bit patterns order like unsigned integers, so two integer
// REDUX steps (high word, then low word among the lanes that hold the winning high word) replace five shuffle rounds of
// NaN-propagating fmax sequences.  Exact.  A NaN (positive quiet pattern) wins, which makes the caller's `norm < eps`
// tests fail instead of silently dropping it.
#ifdef SPECTRAL_CPU_EMU
SP_DEV double sp_warp_max_nonneg(double v) {
  for (int m = 16; m > 0; m >>= 1) { const double o = sp_shfl_xor(v, m); v = (o > v || o != o) ? o : v; }
  return v;
}
#else
SP_DEV double sp_warp_max_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(SP_FULL, hi);
  const unsigned ml = __reduce_max_sync(SP_FULL, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}
#endif
SP_DEV int sp_group_or(int v, int width) {
  for (int m = width >> 1; m > 0; m >>= 1) v = v | sp_shfl_xor_i(v, m);
  return v;
}
SP_DEV int sp_group_max_i(int v, int width) {
  for (int m = width >> 1; m > 0; m >>= 1) {
    int o = sp_shfl_xor_i(v, m);
    v = v > o ? v : o;
  }
  return v;
}

// capacities of the corridor stage
#define SP_REGION_CAP 64   // cubes per region after the split (reference: unbounded std::vector)
#define SP_SELECT_CAP 64   // distinct cubes selected by CollisionCheck
#define SP_MAX_REGIONS 8
#define SP_MAX_KNOTS 256

// Eigen's PartialPivLU inverse of the Bernstein->monomial matrix, column 1, as the shipped
// reference binaries produce it (solve_3d.cc:813): NOT i/5.
#define SP_INVM1_0 0x1.999999999999ap-52
#define SP_INVM1_1 0x1.99999999999a4p-3
#define SP_INVM1_2 0x1.999999999999cp-2
#define SP_INVM1_3 0x1.3333333333334p-1
#define SP_INVM1_4 0x1.999999999999ap-1
#define SP_INVM1_5 0x1.0p+0

struct SpOptionsDev {
  int max_iter, scaling, check_every, adapt_every, polish, polish_refine, polish_rounds, precheck;
  double eps_abs, eps_rel, eps_pinf, rho0, sigma, alpha, adapt_tol, polish_delta, precheck_margin;
};
