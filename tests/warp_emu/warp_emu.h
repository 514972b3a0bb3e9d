// tests/warp_emu/warp_emu.h -- TEST INFRASTRUCTURE: a 32-thread lock-step emulation of the warp
// primitives used by spectral_b200/csrc/*.cuh, so that the kernels' LOGIC can be exercised on a
// machine without a GPU (this build container).  It is a debugging aid for tests/ only: nothing in
// spectral_b200/ or libspectral.so links or loads it, and it is far too slow to be a fallback
// (every shuffle is two pthread barriers).
#pragma once
#include <pthread.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>

#define SP_DEV inline
#define SP_DEV_NOINLINE inline
#define SP_HD inline
#define __host__
#define __device__

struct EmuWarp {
  pthread_barrier_t bar;
  double dslot[32];
  int islot[32];
};
extern thread_local EmuWarp *emu_warp;
extern thread_local int emu_lane;
extern thread_local pthread_barrier_t *emu_cta;  // the CTA barrier of the running emulated block
extern thread_local int emu_warp_id;

inline void emu_sync() { pthread_barrier_wait(&emu_warp->bar); }

inline double emu_exchange_d(double v, int src) {
  emu_warp->dslot[emu_lane] = v;
  emu_sync();
  double r = emu_warp->dslot[src & 31];
  emu_sync();
  return r;
}
inline int emu_exchange_i(int v, int src) {
  emu_warp->islot[emu_lane] = v;
  emu_sync();
  int r = emu_warp->islot[src & 31];
  emu_sync();
  return r;
}
inline double sp_shfl(double v, int src) { return emu_exchange_d(v, src); }
inline int sp_shfl_i(int v, int src) { return emu_exchange_i(v, src); }
inline double sp_shfl_up(double v, int d, int width) {
  int in = emu_lane % width;
  return emu_exchange_d(v, in - d < 0 ? emu_lane : emu_lane - d);
}
inline int sp_shfl_up_i(int v, int d, int width) {
  int in = emu_lane % width;
  return emu_exchange_i(v, in - d < 0 ? emu_lane : emu_lane - d);
}
inline double sp_shfl_down(double v, int d, int width) {
  int in = emu_lane % width;
  return emu_exchange_d(v, in + d >= width ? emu_lane : emu_lane + d);
}
inline double sp_shfl_xor(double v, int m) { return emu_exchange_d(v, emu_lane ^ m); }
inline int sp_shfl_xor_i(int v, int m) { return emu_exchange_i(v, emu_lane ^ m); }
inline unsigned sp_ballot(int pred) {
  emu_warp->islot[emu_lane] = pred ? 1 : 0;
  emu_sync();
  unsigned m = 0;
  for (int i = 0; i < 32; i++) m |= (unsigned)emu_warp->islot[i] << i;
  emu_sync();
  return m;
}
inline int sp_any(int pred) { return sp_ballot(pred) != 0; }
inline int sp_all(int pred) { return sp_ballot(pred) == 0xffffffffu; }
inline void sp_syncwarp() { emu_sync(); }
inline void sp_sync_cta() { pthread_barrier_wait(emu_cta); }
inline int sp_warp_in_cta() { return emu_warp_id; }
inline int sp_popc(unsigned v) { return __builtin_popcount(v); }
inline int sp_ffs(unsigned v) { return __builtin_ffs((int)v); }
// compiled with -ffp-contract=off, so plain operators are IEEE round-to-nearest without FMA
inline double rn_add(double a, double b) { return a + b; }
inline double rn_sub(double a, double b) { return a - b; }
inline double rn_mul(double a, double b) { return a * b; }
inline double rn_div(double a, double b) { return a / b; }
inline void sp_store2(double *p, double a, double b) { p[0] = a; p[1] = b; }
