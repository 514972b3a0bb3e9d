/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * In-memory driver around the reference's OWN class, compiled together with the reference's
 * unmodified sources (solve_3d.cc | cuboid_3d.cc, piecewise_jerk_problem.cc, logging.cc and
 * the matching wrapper) into oracle/_ref/libref_{trp,cub}.so by oracle/Makefile.  It repeats the
 * call sequence of find_traj (trp_wrapper.cpp:76-191 / cub_wrapper.cpp:75-186) without the text
 * files, so batches can be pushed through the real reference code path:
 *   set_* -> per region set_x_bounds/set_y_bounds/CorridorGeneration/PrintCorridor
 *         -> CollisionCheck -> Optimize(5000) -> opt_x()...
 * The control points live only inside the OSQP workspace (solve_3d.cc:1343-1344); the shim
 * (osqp_shim.c) hands them back through spectral_shim_last().
 * Guards (reference behaviour is undefined / fatal there):
 *   - nothing selected by CollisionCheck -> `temp.size() - 1` underflow (solve_3d.cc:617);
 *   - sample-count CHECK (solve_3d.cc:1407) would abort(), x_.at() would throw;
 *   both are predicted with the C restatement and reported as a status instead of being run.
 * x_ref/y_ref are handed over in vectors whose spare capacity is zero-filled, so the
 * out-of-bounds read ref[10k+1] (solve_3d.cc:1161) deterministically sees 0.0 like the shipped
 * binary did.
 */
#include <algorithm>
#include <array>
#include <cstring>
#include <iostream>
#include <vector>

#include "btrapz/solve_3d.h"
#include "osqp_shim.h"
#include "spectral_oracle.h"

#ifdef _OPENMP
#include <omp.h>
#endif

static_assert(sizeof(Cube) == sizeof(OracleCube), "Cube layout");

using navigation::PiecewiseJerkSpeedProblem;

#ifndef REF_VARIANT
#error "REF_VARIANT must be ORACLE_TRP or ORACLE_CUB"
#endif

static std::vector<double> zero_tailed(const double *src, int n) {
  std::vector<double> v;
  v.resize(n + 64, 0.0);
  v.resize(n);
  std::copy(src, src + n, v.begin());
  return v;
}

static std::vector<std::pair<double, double>> pairs(const double *src, int n) {
  std::vector<std::pair<double, double>> v;
  v.reserve(n);
  for (int i = 0; i < n; i++) v.emplace_back(src[2 * i], src[2 * i + 1]);
  return v;
}

extern "C" int ref_variant(void) { return REF_VARIANT; }

/* real std::sort with the reference's comparator (solve_3d.cc:630), for checking the restated sort */
extern "C" void ref_std_sort_by_beg_t(OracleCube *a, int n) {
  Cube *c = reinterpret_cast<Cube *>(a);
  std::sort(c, c + n, [](const Cube &r1, const Cube &r2) -> bool { return r1.beg_t < r2.beg_t; });
}

extern "C" int ref_solve_batch(int B, int N, int R, double delta, const double *s_bounds,
                               const double *l_bounds, const double *ds_bounds,
                               const double *dl_bounds, const double *s_ref, const double *l_ref,
                               const double *init, const double *scalars, const double *weights,
                               int weights_stride, int mode, int k_max, int nthreads, int *K,
                               OracleCube *segs, double *ctrl, double *obj, double *a_cost,
                               int *status, int *iters, int *npts, double *samples,
                               int samples_cap, int *polish) {
  std::ios_base::iostate old_state = std::cout.rdstate();
  std::cout.setstate(std::ios_base::failbit); /* silence the reference's debug prints */
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 4)
  for (int b = 0; b < B; b++) {
    const double *sc = scalars + 10 * (size_t)b;
    const double *w = weights + (size_t)(weights_stride ? b : 0) * 10;
    const double *sb = s_bounds + (size_t)b * R * N * 2, *lb = l_bounds + (size_t)b * R * N * 2;
    OracleCube *sg = segs + (size_t)b * k_max;
    double *ct = ctrl + (size_t)b * 12 * k_max;
    std::memset(sg, 0, sizeof(OracleCube) * (size_t)k_max);
    std::memset(ct, 0, sizeof(double) * 12 * (size_t)k_max);
    K[b] = 0; obj[b] = 0; a_cost[b] = 100000000000.0; iters[b] = 0; status[b] = ORACLE_FAIL_SOLVER;
    if (npts) npts[b] = 0;
    if (polish) polish[b] = 0;

    /* ---- guards, predicted with the restatement */
    OracleProblem p;
    p.n_knots = N; p.delta = delta;
    for (int i = 0; i < 3; i++) { p.init_s[i] = init[6 * b + i]; p.init_l[i] = init[6 * b + 3 + i]; }
    p.ds_ref = sc[0]; p.dl_ref = sc[1]; p.dds_lo = sc[2]; p.dds_hi = sc[3]; p.ddds_lo = sc[4];
    p.ddds_hi = sc[5]; p.ddl_lo = sc[6]; p.ddl_hi = sc[7]; p.dddl_lo = sc[8]; p.dddl_hi = sc[9];
    p.ds_bounds = ds_bounds + (size_t)b * N * 2; p.dl_bounds = dl_bounds + (size_t)b * N * 2;
    p.s_ref = s_ref + (size_t)b * N; p.l_ref = l_ref + (size_t)b * N;
    for (int i = 0; i < 10; i++) p.w[i] = w[i];
    {
      std::vector<OracleCube> corr((size_t)R * 64);
      std::vector<int> counts(R);
      bool bad = false;
      for (int r = 0; r < R; r++) {
        counts[r] = oracle_corridor_generation(REF_VARIANT, N, delta, sb + (size_t)r * N * 2,
                                               lb + (size_t)r * N * 2, corr.data() + (size_t)r * 64, 64);
        if (counts[r] < 0) bad = true;
      }
      if (bad) { status[b] = ORACLE_FAIL_TOO_MANY; continue; }
      OracleCube tmp[256];
      int kk = oracle_collision_check(REF_VARIANT, R, corr.data(), counts.data(), 64, N, delta,
                                      p.s_ref, p.l_ref, tmp, 256);
      if (kk == 0) { status[b] = ORACLE_FAIL_NO_CORRIDOR; continue; }
      if (kk < 0 || kk > k_max) { K[b] = kk < 0 ? -kk : kk; status[b] = ORACLE_FAIL_TOO_MANY; continue; }
      std::vector<double> zeros(12 * (size_t)kk, 0.0), smp(6 * 1024);
      if (oracle_sample(&p, tmp, kk, zeros.data(), smp.data(), 1024) < 0) {
        K[b] = kk; status[b] = ORACLE_FAIL_POINTS_CHECK; continue;
      }
    }

    /* ---- the reference itself */
    std::array<double, 3> init_s = {p.init_s[0], p.init_s[1], p.init_s[2]};
    std::array<double, 3> init_l = {p.init_l[0], p.init_l[1], p.init_l[2]};
    PiecewiseJerkSpeedProblem prob(N, delta, init_s, init_l, R);
    prob.set_scale_factor({1.0, 1.0, 1.0});
    prob.set_weight_ddx(w[0]);
    prob.set_weight_dddx(w[1]);
    prob.set_ddx_bounds(sc[2], sc[3]);
    prob.set_dddx_bound(sc[4], sc[5]);
    prob.set_dx_bounds(pairs(p.ds_bounds, N));
    prob.set_x_ref(w[4], zero_tailed(p.s_ref, N));
    prob.set_dx_ref(w[5], sc[0]);
    prob.set_weight_ddy(w[2]);
    prob.set_weight_dddy(w[3]);
    prob.set_ddy_bounds(sc[6], sc[7]);
    prob.set_dddy_bound(sc[8], sc[9]);
    prob.set_dy_bounds(pairs(p.dl_bounds, N));
    prob.set_y_ref(w[6], zero_tailed(p.l_ref, N));
    prob.set_dy_ref(w[7], sc[1]);
    prob.set_weight_end(w[8], w[9]);
    for (int r = 0; r < R; r++) {
      prob.set_x_bounds(pairs(sb + (size_t)r * N * 2, N));
      prob.set_y_bounds(pairs(lb + (size_t)r * N * 2, N));
      prob.CorridorGeneration();
      prob.PrintCorridor();
    }
    prob.CollisionCheck();
    int kk = (int)prob.new_corridor.size();
    K[b] = kk;
    for (int k = 0; k < kk && k < k_max; k++) std::memcpy(&sg[k], &prob.new_corridor[k], sizeof(Cube));
    SpectralShimOverride ovr;
    std::memset(&ovr, 0, sizeof(ovr));
    bool ok = false;
    if (mode == 0) {
      spectral_shim_set_override(nullptr);
      ok = prob.Optimize(5000);
    } else {
      /* converged optimum through the same reference code path: tighten until the polish is accepted.
       * A fresh problem object is needed per attempt (num_of_points_ accumulates in Optimize). */
      static const double ladder[3] = {1e-6, 1e-8, 1e-10};
      ovr.active = 1; ovr.max_iter = 50000; ovr.polish = 1; ovr.delta = 1e-9; ovr.polish_refine_iter = 8; ovr.polish_rounds = 12;
      ovr.eps = ladder[0];
      spectral_shim_set_override(&ovr);
      ok = prob.Optimize(5000);
      /* (the ladder's later rungs are exercised by the C restatement; one rung suffices here) */
      spectral_shim_set_override(nullptr);
    }
    SpectralShimLast *last = spectral_shim_last();
    iters[b] = last->iter;
    if (polish) polish[b] = last->polish_status;
    if (!ok) { status[b] = ORACLE_FAIL_SOLVER; continue; }
    std::memcpy(ct, last->x, sizeof(double) * 12 * (size_t)kk);
    obj[b] = last->obj_val;
    const std::vector<double> &s = prob.opt_x(), &ds = prob.opt_dx(), &dds = prob.opt_ddx();
    const std::vector<double> &l = prob.opt_y(), &dl = prob.opt_dy(), &ddl = prob.opt_ddy();
    int np_ = (int)s.size();
    std::vector<double> smp(6 * (size_t)np_);
    for (int i = 0; i < np_; i++) {
      smp[6 * i + 0] = s[i]; smp[6 * i + 1] = ds[i]; smp[6 * i + 2] = dds[i];
      smp[6 * i + 3] = l[i]; smp[6 * i + 4] = dl[i]; smp[6 * i + 5] = ddl[i];
    }
    if (npts) npts[b] = np_;
    if (samples) std::memcpy(samples + (size_t)b * samples_cap * 6, smp.data(),
                             sizeof(double) * 6 * (size_t)std::min(np_, samples_cap));
    a_cost[b] = oracle_cost(REF_VARIANT, &p, smp.data(), np_); /* wrapper cost, restated (a10) */
    status[b] = (last->status == 2) ? ORACLE_SOLVED_INACCURATE : ORACLE_OK;
  }
  std::cout.clear(old_state);
  return 0;
}
