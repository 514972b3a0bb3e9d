#!/usr/bin/env python
"""oracle/run_shipped.py -- TEST INFRASTRUCTURE (golden generation), not product code.

Runs the reference's SHIPPED binaries (/root/reference/src/libtrp.so, libcub.so) in a
child process behind oracle/_ref/libosqp.so (capture shim) and oracle/_ref/libpathredirect.so
(maps the hard-coded /home/srujan_d/... paths of trp_wrapper.cpp:23,288 into a scratch dir).
Only usable in the build container (needs /root/reference); its outputs are committed under
tests/golden/ by oracle/gen_golden.py.
"""
import ctypes
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
INPUT_NAME = {"trp": "c_road_s1_2.txt", "cub": "c_road_s1_3.txt"}
OUTPUT_FMT = {"trp": "s1_slt_3d_%d.txt", "cub": "s1_cub_3d_%d.txt"}

CHILD = r"""
import ctypes, sys
class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "s_acc_weight","s_jerk_weight","l_acc_weight","l_jerk_weight","weight_s_ref",
        "weight_ds_ref","weight_l_ref","weight_dl_ref","weight_end_s","weight_end_l")] + [("iteration", ctypes.c_int)]
lib = ctypes.CDLL(sys.argv[1])
lib.find_traj.argtypes = (ctypes.POINTER(Params),)
lib.find_traj.restype = ctypes.c_double
w = [float(v) for v in sys.argv[2:12]]
p = Params(*w, int(sys.argv[12]))
r = lib.find_traj(ctypes.byref(p))
sys.stdout.flush()
import os
os.write(2, ("RETVAL %r\n" % r).encode())
"""


def run_shipped(variant, fixture, weights, iteration=31, solution=None, lib=None, keep_stdout=False):
    """Returns dict(qp_text=..., retval=..., out_text=... or None, stdout=...)."""
    lib = lib or os.path.join(REF_SRC, "lib%s.so" % variant)
    io_dir = tempfile.mkdtemp(prefix="spectral_io_")
    try:
        shutil.copy(fixture, os.path.join(io_dir, INPUT_NAME[variant]))
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = os.path.join(HERE, "_ref") + ":" + env.get("LD_LIBRARY_PATH", "")
        env["LD_PRELOAD"] = os.path.join(HERE, "_ref", "libpathredirect.so")
        env["SPECTRAL_IO_DIR"] = io_dir
        env["SPECTRAL_QP_DUMP"] = os.path.join(io_dir, "qp.txt")
        if solution is not None:
            sol_path = os.path.join(io_dir, "sol.bin")
            import numpy as np
            np.asarray(solution, dtype=np.float64).tofile(sol_path)
            env["SPECTRAL_QP_SOLUTION"] = sol_path
        else:
            env.pop("SPECTRAL_QP_SOLUTION", None)
        args = [sys.executable, "-c", CHILD, lib] + ["%r" % float(w) for w in weights] + [str(iteration)]
        pr = subprocess.run(args, env=env, capture_output=True, timeout=120)
        err = pr.stderr.decode(errors="replace")
        retval = None
        for line in err.splitlines():
            if "RETVAL " in line:
                retval = float(line.split("RETVAL ")[1].split()[0])
        out = {"retval": retval, "returncode": pr.returncode, "stderr": err}
        if keep_stdout:
            out["stdout"] = pr.stdout.decode(errors="replace")
        qp = os.path.join(io_dir, "qp.txt")
        out["qp_text"] = open(qp).read() if os.path.exists(qp) else None
        of = os.path.join(io_dir, OUTPUT_FMT[variant] % iteration)
        out["out_text"] = open(of).read() if os.path.exists(of) else None
        return out
    finally:
        shutil.rmtree(io_dir, ignore_errors=True)


def parse_qp(text):
    """Parse the capture shim's dump into a dict of numpy arrays."""
    import numpy as np
    toks = text.split("\n")
    out = {}
    i = 0
    while i < len(toks):
        line = toks[i].strip()
        i += 1
        if not line:
            continue
        parts = line.split()
        if parts[0] in ("n", "m") and len(parts) == 2:
            out[parts[0]] = int(parts[1])
        elif parts[0] == "settings":
            s = {}
            for k in range(1, len(parts), 2):
                s[parts[k]] = float(parts[k + 1])
            out["settings"] = s
        else:
            name, cnt = parts[0], int(parts[1])
            vals = toks[i:i + cnt]
            i += cnt
            if name.endswith("_p") or name.endswith("_i"):
                out[name] = np.array([int(v) for v in vals], dtype=np.int64)
            else:
                out[name] = np.array([float(v) for v in vals], dtype=np.float64)
    return out


GOLDEN_W_TRP = (35.73, 41.61, 25.57, 41.59, 0.12, 10.04, 0.0, 0.0, 7.27, 32.13)
GOLDEN_W_CUB = (35.73, 41.61, 25.57, 41.59, 0.12, 10.04, 0.0, 0.0, 7.27, 0.0)

if __name__ == "__main__":
    variant, fixture = sys.argv[1], sys.argv[2]
    w = GOLDEN_W_TRP if variant == "trp" else GOLDEN_W_CUB
    r = run_shipped(variant, fixture, w, keep_stdout=True)
    print("retval", r["retval"], "rc", r["returncode"])
    qp = parse_qp(r["qp_text"]) if r["qp_text"] else None
    if qp:
        print("n", qp["n"], "m", qp["m"], "nnzP", len(qp["P_x"]), "nnzA", len(qp["A_x"]))
        print(qp["settings"])
    print(r["stdout"][-1500:])
    print(r["stderr"][-500:])
