"""oracle/pyoracle.py -- TEST INFRASTRUCTURE: ctypes bindings of the CPU oracle.

Loads oracle/liboracle.so (plain-C restatement) and, when present, oracle/_ref/libref_{trp,cub}.so
(the reference's own sources recompiled).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module; the product never does.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TRP, CUB = 0, 1
VARIANT_ID = {"trp": TRP, "cub": CUB}

STATUS_OK, STATUS_INACCURATE, STATUS_NO_CORRIDOR, STATUS_SOLVER, STATUS_POINTS, STATUS_TOO_MANY = range(6)

CUBE_DTYPE = np.dtype([
    ("beg_t", np.int32), ("end_t", np.int32), ("t", np.float64), ("t_dif", np.float64),
    ("beg_l", np.float64), ("end_l", np.float64), ("upp_skew", np.float64), ("upp_bias", np.float64),
    ("down_skew", np.float64), ("down_bias", np.float64), ("l_upp_skew", np.float64),
    ("l_upp_bias", np.float64), ("l_down_skew", np.float64), ("l_down_bias", np.float64),
    ("merge", np.uint8), ("split", np.uint8), ("_pad", np.uint8, (2,)), ("count", np.int32)])
assert CUBE_DTYPE.itemsize == 112

CUBE_FIELDS = ("beg_t", "end_t", "t", "beg_l", "end_l", "upp_skew", "upp_bias", "down_skew", "down_bias",
               "l_upp_skew", "l_upp_bias", "l_down_skew", "l_down_bias")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


_libs = {}


def load(kind="port", variant="trp"):
    """kind: 'port' (liboracle.so) or 'reference' (oracle/_ref/libref_<variant>.so)."""
    key = (kind, variant if kind == "reference" else "")
    if key in _libs:
        return _libs[key]
    path = os.path.join(HERE, "liboracle.so") if kind == "port" else os.path.join(HERE, "_ref", "libref_%s.so" % variant)
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (run `make -C oracle`)")
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    _libs[key] = lib
    return lib


def have_reference():
    return all(os.path.exists(os.path.join(HERE, "_ref", "libref_%s.so" % v)) for v in ("trp", "cub"))


def solve_batch(variant, batch, weights, mode=0, k_max=32, nthreads=1, kind="port", samples_cap=160):
    """Run the CPU oracle over a ScenarioBatch.  weights: [10] or [B,10].  Returns a dict of arrays."""
    B, N, R = batch.batch, batch.n_knots, batch.n_regions
    # the sample buffer must hold every sample the path can produce (1 + 10 per segment): a buffer that is too small
    # would be reported as the reference's CHECK failure by oracle_sample()
    samples_cap = max(samples_cap, 10 * k_max + 8)
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
    stride = 0 if w.ndim == 1 else 1
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in batch.arrays()]
    K = np.zeros(B, np.int32)
    segs = np.zeros((B, k_max), CUBE_DTYPE)
    ctrl = np.zeros((B, 12 * k_max))
    obj = np.zeros(B)
    a_cost = np.zeros(B)
    status = np.zeros(B, np.int32)
    iters = np.zeros(B, np.int32)
    npts = np.zeros(B, np.int32)
    samples = np.zeros((B, samples_cap, 6))
    polish = np.zeros(B, np.int32)
    lib = load(kind, variant)
    vid = VARIANT_ID[variant]
    common = [ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(R), ctypes.c_double(batch.delta_t)] + \
        [_d(a) for a in arrs] + [_d(w), ctypes.c_int(stride), ctypes.c_int(mode), ctypes.c_int(k_max),
                                 ctypes.c_int(nthreads), _i(K), segs.ctypes.data_as(ctypes.c_void_p), _d(ctrl),
                                 _d(obj), _d(a_cost), _i(status), _i(iters), _i(npts), _d(samples),
                                 ctypes.c_int(samples_cap), _i(polish)]
    if kind == "port":
        lib.oracle_solve_batch(ctypes.c_int(vid), *common)
    else:
        assert lib.ref_variant() == vid
        lib.ref_solve_batch(*common)
    return dict(K=K, segs=segs, ctrl=ctrl, obj=obj, a_cost=a_cost, status=status, iters=iters, npts=npts,
                samples=samples, polish=polish)


class _QP(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("m", ctypes.c_int),
                ("P_p", ctypes.POINTER(ctypes.c_longlong)), ("P_i", ctypes.POINTER(ctypes.c_longlong)), ("P_x", _dp),
                ("A_p", ctypes.POINTER(ctypes.c_longlong)), ("A_i", ctypes.POINTER(ctypes.c_longlong)), ("A_x", _dp),
                ("q", _dp), ("l", _dp), ("u", _dp)]


class _Problem(ctypes.Structure):
    _fields_ = [("n_knots", ctypes.c_int), ("delta", ctypes.c_double), ("init_s", ctypes.c_double * 3),
                ("init_l", ctypes.c_double * 3), ("ds_ref", ctypes.c_double), ("dl_ref", ctypes.c_double),
                ("dds_lo", ctypes.c_double), ("dds_hi", ctypes.c_double), ("ddds_lo", ctypes.c_double),
                ("ddds_hi", ctypes.c_double), ("ddl_lo", ctypes.c_double), ("ddl_hi", ctypes.c_double),
                ("dddl_lo", ctypes.c_double), ("dddl_hi", ctypes.c_double), ("ds_bounds", _dp), ("dl_bounds", _dp),
                ("s_ref", _dp), ("l_ref", _dp), ("w", ctypes.c_double * 10)]


def formulate(variant, sc, weights, segs):
    """Assemble the QP (a5-a8) for one Scenario and a given segment list -> dict like run_shipped.parse_qp."""
    lib = load("port")
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (sc.ds_bounds, sc.dl_bounds, sc.s_ref, sc.l_ref)]
    p = _Problem()
    p.n_knots = sc.n_knots
    p.delta = sc.delta_t
    for i in range(3):
        p.init_s[i] = sc.init_s[i]
        p.init_l[i] = sc.init_l[i]
    (p.ds_ref, p.dl_ref, p.dds_lo, p.dds_hi, p.ddds_lo, p.ddds_hi, p.ddl_lo, p.ddl_hi, p.dddl_lo,
     p.dddl_hi) = [float(v) for v in sc.scalars]
    p.ds_bounds, p.dl_bounds, p.s_ref, p.l_ref = [_d(a) for a in keep]
    for i in range(10):
        p.w[i] = float(weights[i])
    segs = np.ascontiguousarray(segs)
    K = len(segs)
    qp = _QP()
    rc = lib.oracle_formulate(ctypes.c_int(VARIANT_ID[variant]), ctypes.byref(p),
                              segs.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(K), ctypes.byref(qp))
    if rc != 0:
        raise RuntimeError("oracle_formulate failed")
    n, m = qp.n, qp.m
    P_p = np.ctypeslib.as_array(qp.P_p, (n + 1,)).copy()
    A_p = np.ctypeslib.as_array(qp.A_p, (n + 1,)).copy()
    out = dict(n=n, m=m, P_p=P_p, P_i=np.ctypeslib.as_array(qp.P_i, (P_p[-1],)).copy(),
               P_x=np.ctypeslib.as_array(qp.P_x, (P_p[-1],)).copy(), A_p=A_p,
               A_i=np.ctypeslib.as_array(qp.A_i, (A_p[-1],)).copy(),
               A_x=np.ctypeslib.as_array(qp.A_x, (A_p[-1],)).copy(),
               q=np.ctypeslib.as_array(qp.q, (n,)).copy(), l=np.ctypeslib.as_array(qp.l, (m,)).copy(),
               u=np.ctypeslib.as_array(qp.u, (m,)).copy())
    lib.oracle_qp_free(ctypes.byref(qp))
    return out


def std_sort_check(beg_t, kind="port", variant="trp"):
    """Sort cubes keyed by beg_t (count field = original index) with the restated / the real std::sort."""
    n = len(beg_t)
    cubes = np.zeros(n, CUBE_DTYPE)
    cubes["beg_t"] = beg_t
    cubes["count"] = np.arange(n)
    lib = load(kind, variant)
    fn = lib.oracle_std_sort_by_beg_t if kind == "port" else lib.ref_std_sort_by_beg_t
    fn(cubes.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n))
    return cubes["count"].copy()
