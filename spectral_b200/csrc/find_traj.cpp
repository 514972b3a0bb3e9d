// spectral_b200/csrc/find_traj.cpp -- the reference's plugin entry point on top of libspectral.so.
//
// Builds libtrp.so (-DSPECTRAL_VARIANT=0) and libcub.so (-DSPECTRAL_VARIANT=1) exporting exactly
//     extern "C" double find_traj(Params *p)
// as /root/reference/src/trp_wrapper.cpp:16-306 and cub_wrapper.cpp:16-285 do, so that the reference's
// ctypes bindings (src/trp_wrapper.py:45-54, src/cub_wrapper.py) load them unchanged:
//   - hidden input  <dir>/c_road_s1_2.txt (trp) | <dir>/c_road_s1_3.txt (cub)    trp_wrapper.cpp:23
//   - hidden output <dir>/s1_slt_3d_<it>.txt   | <dir>/s1_cub_3d_<it>.txt, "t s l ds dl dds ddl",
//     std::fixed, setprecision(3), one row per sample                              trp_wrapper.cpp:288-301
//   - return value  a_cost, or 100000000000 (and no file) when the optimisation fails  :195-200,304
// <dir> is the reference's literal /home/srujan_d/RISS/code/btrapz/src unless $SPECTRAL_IO_DIR is set.
// Deviations, all where the reference is undefined: an unreadable input file returns the
// failure sentinel instead of computing on garbage; cub's uninitialised l_cost starts at 0; trp's
// end-term reads the last available sample instead of out of bounds.
// The numerical work happens on the GPU; there is no CPU path.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/spectral.h"

#ifndef SPECTRAL_VARIANT
#error "define SPECTRAL_VARIANT (0 = trp, 1 = cub)"
#endif

namespace {
const char *kDefaultDir = "/home/srujan_d/RISS/code/btrapz/src";
// One lazily created handle per process, guarded by a mutex (the reference's find_traj is single-threaded and not
// re-entrant w.r.t. its fixed file names either, trp_wrapper.cpp:23,288) and destroyed at unload.
std::mutex g_mu;
struct HandleOwner {
  spectral_handle_t *h = nullptr;
  int nmax = 0, rmax = 0;
  ~HandleOwner() { if (h) spectral_destroy(h); }
} g_own;

std::string io_dir() {
  const char *d = getenv("SPECTRAL_IO_DIR");
  return d ? std::string(d) : std::string(kDefaultDir);
}
bool verbose() { const char *v = getenv("SPECTRAL_VERBOSE"); return v && v[0] == '1'; }
}  // namespace

extern "C" double find_traj(SpectralParams *p) {
  const double kFail = SPECTRAL_FAIL_COST;
  std::lock_guard<std::mutex> lock(g_mu);
  spectral_handle_t *&g_handle = g_own.h;
  int &g_nmax = g_own.nmax, &g_rmax = g_own.rmax;
  const std::string in_path = io_dir() + (SPECTRAL_VARIANT == SPECTRAL_TRP ? "/c_road_s1_2.txt" : "/c_road_s1_3.txt");
  std::ifstream ifs(in_path);
  if (!ifs.is_open()) {
    std::cerr << "find_traj: cannot read " << in_path << std::endl;
    return kFail;
  }
  int N = 0, R = 0;
  double delta_t = 0;
  double init[6], scalars[10];
  ifs >> N >> delta_t;                                  // trp_wrapper.cpp:39
  ifs >> init[0] >> init[1] >> init[2];                 // :40
  ifs >> init[3] >> init[4] >> init[5];                 // :41
  ifs >> R;                                             // :49
  ifs >> scalars[0] >> scalars[1];                      // ds_ref dl_ref :59
  for (int i = 2; i < 10; i++) ifs >> scalars[i];       // dd/ddd bounds :61-64
  if (!ifs || N < 3 || R < 1) {
    std::cerr << "find_traj: malformed header in " << in_path << std::endl;
    return kFail;
  }
  if (N > 256 || R > 8) {  // capacity of the corridor kernel's shared-memory slabs (SP_MAX_KNOTS / SP_MAX_REGIONS)
    std::cerr << "find_traj: " << in_path << " has " << N << " knots / " << R << " regions; this build supports at most 256 / 8"
              << std::endl;
    return kFail;
  }
  std::vector<double> sb((size_t)R * N * 2), lb((size_t)R * N * 2), dsb((size_t)N * 2), dlb((size_t)N * 2), sref(N), lref(N);
  for (int r = 0; r < R; r++) {                          // :77-100
    for (int i = 0; i < 2 * N; i++) ifs >> sb[(size_t)r * N * 2 + i];
    for (int i = 0; i < 2 * N; i++) ifs >> lb[(size_t)r * N * 2 + i];
  }
  for (int i = 0; i < 2 * N; i++) ifs >> dsb[i];         // :102-106
  for (int i = 0; i < 2 * N; i++) ifs >> dlb[i];         // :108-113
  for (int i = 0; i < N; i++) ifs >> sref[i];            // :116-122
  for (int i = 0; i < N; i++) ifs >> lref[i];            // :125-131
  if (!ifs) {
    std::cerr << "find_traj: truncated input " << in_path << std::endl;
    return kFail;
  }
  // (the kappa columns that follow are read and never used by the reference, :134-144)

  if (!g_handle || N > g_nmax || R > g_rmax) {
    if (g_handle) spectral_destroy(g_handle);
    g_handle = nullptr;
    g_nmax = N > 128 ? N : 128;
    g_rmax = R > 8 ? R : 8;
    const char *dev = getenv("SPECTRAL_DEVICE");
    if (spectral_create(dev ? atoi(dev) : 0, 1, g_nmax, g_rmax, 32, &g_handle) != SPECTRAL_SUCCESS) {
      std::cerr << "find_traj: " << (g_handle ? spectral_last_error(g_handle) : "no CUDA device") << std::endl;
      if (g_handle) spectral_destroy(g_handle);
      g_handle = nullptr;
      return kFail;
    }
  }
  const double w[10] = {p->s_acc_weight, p->s_jerk_weight, p->l_acc_weight, p->l_jerk_weight, p->weight_s_ref,
                        p->weight_ds_ref, p->weight_l_ref, p->weight_dl_ref, p->weight_end_s, p->weight_end_l};
  SpectralInputs in;
  in.s_bounds = sb.data(); in.l_bounds = lb.data(); in.ds_bounds = dsb.data(); in.dl_bounds = dlb.data();
  in.s_ref = sref.data(); in.l_ref = lref.data(); in.init = init; in.scalars = scalars; in.weights = w; in.weights_stride = 0;
  const int cap = 512;
  std::vector<double> samples((size_t)cap * 6), ctrl(12 * 32);
  std::vector<SpectralCube> segs(32);
  int K = 0, status = 0, iters = 0, flags = 0, npts = 0;
  double obj = 0, a_cost = kFail;
  SpectralOutputs out;
  out.K = &K; out.segs = segs.data(); out.ctrl = ctrl.data(); out.obj = &obj; out.a_cost = &a_cost; out.status = &status;
  out.iters = &iters; out.flags = &flags; out.npts = &npts; out.samples = samples.data(); out.samples_cap = cap; out.lu = nullptr;
  // Options: the library defaults = the reference's OSQP settings (solve_3d.cc:1235-1243,1446-1462) plus the polish
  // step, which returns the exact optimum of the QP instead of OSQP's eps = 1e-5 iterate (the reference sets polish = 0,
  // :1243).  SPECTRAL_POLISH=0 switches it off: the output is then the raw ADMM iterate like the reference's.
  SpectralOptions opt;
  spectral_default_options(&opt);
  if (const char *pol = getenv("SPECTRAL_POLISH")) opt.polish = pol[0] != '0';
  if (spectral_solve_batch(g_handle, SPECTRAL_VARIANT, 1, N, R, delta_t, &in, &opt, &out) != SPECTRAL_SUCCESS) {
    std::cerr << "find_traj: " << spectral_last_error(g_handle) << std::endl;
    return kFail;
  }
  if (verbose()) {
    std::cout << "\n\n new corridors are \n\n";  // solve_3d.cc:676,708-709
    for (int k = 0; k < K; k++)
      std::cout << segs[k].beg_t << " " << segs[k].end_t << " " << segs[k].down_bias << " " << segs[k].upp_bias << " "
                << segs[k].down_bias + segs[k].down_skew * delta_t << " " << segs[k].upp_skew * delta_t + segs[k].upp_bias
                << " " << segs[k].beg_l << " " << segs[k].end_l << "\n";
    std::cout << "status " << status << " iters " << iters << " flags " << flags << " obj " << obj << "\n";
  }
  if (status != SPECTRAL_SOLVED && status != SPECTRAL_SOLVED_INACCURATE) {
    std::cerr << "Piecewise jerk speed optimizer failed!" << std::endl;  // trp_wrapper.cpp:197-199
    return kFail;
  }
  if (verbose()) {
    // the reference's remaining print blocks (trp_wrapper.cpp:116-132 refs, :217-239 s part, :243-280 l part): the same
    // lines in the same order, with the cost terms accumulated on the host from the samples exactly as the reference
    // does (the returned a_cost is the device's; the trp end term reads the last available sample)
    std::cout << "\nx ref:\n";
    for (int i = 0; i < N; i++) std::cout << sref[i] << " ";
    std::cout << "\ny ref:\n";
    for (int i = 0; i < N; i++) std::cout << lref[i] << " ";
    std::cout << "\n\n";
    const int np = npts < cap ? npts : cap;
    for (int axis = 0; axis < 2; axis++) {
      const int o = 3 * axis;
      const double w_ref = w[axis == 0 ? 4 : 6], w_dref = w[axis == 0 ? 5 : 7], w_acc = w[axis == 0 ? 0 : 2], w_jerk = w[axis == 0 ? 1 : 3];
      const std::vector<double> &ref = axis == 0 ? sref : lref;
      double cost = 0.0, mmax_a = 0.0;
      if (axis == 1) std::cout << "\n\n\nL part of trajectory: \n\n\n\n\nprinting L part: \n\nl size is\t" << np << "\n\n";
      for (int i = 0; i < np; ++i) {
        const double *q = &samples[(size_t)i * 6 + o];
        const double dd_prev = i == 0 ? q[2] : samples[(size_t)(i - 1) * 6 + o + 2];
        const double dd_next = (i == 0 && np > 1) ? samples[(size_t)6 + o + 2] : q[2];
        const double ddd = (i == 0) ? (dd_next - q[2]) / delta_t : (q[2] - dd_prev) / delta_t;
        const double r = i < N ? ref[i] : 0.0;
        cost += w_ref * (q[0] - r) * (q[0] - r) * delta_t + w_dref * q[1] * q[1] * delta_t + w_acc * q[2] * q[2] * delta_t +
                w_jerk * ddd * ddd * delta_t;
        mmax_a = std::max(mmax_a, std::fabs(q[2]));
        std::cout << std::fixed << std::setprecision(3) << "For t[" << i * delta_t << "], opt = " << q[0] << ", " << q[1] << ", " << q[2]
                  << ", " << ddd << std::endl;
      }
      if (axis == 1 && np > 0 && SPECTRAL_VARIANT == SPECTRAL_TRP) {
        const double e = samples[(size_t)(np - 1) * 6 + 3] - lref[np - 1 < N ? np - 1 : N - 1];
        cost += w[9] * e * e * delta_t;
      }
      std::cout << (axis == 0 ? "\ns_cost is \t" : "\nl_cost is \t") << cost << "\n";
      std::cout << "mmax_a " << mmax_a << std::endl;
    }
    std::cout << "a_cost " << a_cost << std::endl;
  }
  std::string file = io_dir() + (SPECTRAL_VARIANT == SPECTRAL_TRP ? "/s1_slt_3d_" : "/s1_cub_3d_");
  file += std::to_string(p->iteration) + ".txt";
  std::ofstream ofs(file);
  const int n_out = npts < cap ? npts : cap;
  for (int i = 0; i < n_out; ++i) {  // :298-301
    const double *s = &samples[(size_t)i * 6];
    ofs << std::fixed << std::setprecision(3) << i * delta_t << " " << s[0] << " " << s[3] << " " << s[1] << " " << s[4]
        << " " << s[2] << " " << s[5] << std::endl;
  }
  return a_cost;
}
