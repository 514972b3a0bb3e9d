/* oracle/osqp_shim.h -- TEST INFRASTRUCTURE, not product code (see osqp_shim.c). */
#ifndef SPECTRAL_ORACLE_OSQP_SHIM_H
#define SPECTRAL_ORACLE_OSQP_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
#define SPECTRAL_SHIM_MAX_N 768
typedef struct {
  int n, m, status, iter, polish_status, rho_updates;
  double obj_val, pri_res, dua_res;
  double x[SPECTRAL_SHIM_MAX_N];
} SpectralShimLast;
typedef struct {
  int active;           /* 0: use exactly the settings the reference passed */
  double eps;           /* eps_abs = eps_rel */
  long long max_iter;
  long long polish;
  double delta;
  long long polish_refine_iter;
  long long polish_rounds;
} SpectralShimOverride;
SpectralShimLast *spectral_shim_last(void);
void spectral_shim_set_override(const SpectralShimOverride *o);
#ifdef __cplusplus
}
#endif
#endif
