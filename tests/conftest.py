"""pytest configuration: the `gpu` marker, sys.path, and one-time builds of the CPU-side test libraries."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _make(path):
    r = subprocess.run(["make", "-C", os.path.join(ROOT, path)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("make -C %s failed:\n%s\n%s" % (path, r.stdout[-3000:], r.stderr[-3000:]))


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The oracle (checker) and the kernel-logic emulator are plain gcc/g++ builds (seconds).  The CUDA
    libraries are built by __graft_entry__.build(); on the GPU box the prebuilt ones travel with the repo."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        _make("oracle")
    if not os.path.exists(os.path.join(ROOT, "tests", "warp_emu", "libwarp_emu.so")):
        _make("tests/warp_emu")
    if not os.path.exists(os.path.join(ROOT, "spectral_b200", "lib", "libspectral.so")):
        _make("spectral_b200/csrc")
    yield


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run via gpurun / the driver)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """One JSON line per parity check of the session (tests/helpers.assert_batch_parity) plus one aggregate line, in the
    pytest output itself, so that the driver's log keeps the measured parity statistics: how many scenarios were compared
    and how (KKT-verified optimum vs ADMM iterate), iteration-count agreement, class mismatches, exceptions."""
    try:
        import json
        import helpers
    except Exception:
        return
    log = helpers.PARITY_LOG
    if not log:
        return
    tr = terminalreporter
    tr.write_line("")
    for s in log:
        tr.write_line(json.dumps({"parity_stats": s}))
    tot = lambda k: int(sum((s.get(k) or 0) for s in log))  # noqa: E731
    it_c = tot("iters_compared")
    it_eq = sum((s.get("iters_equal_frac") or 0.0) * (s.get("iters_compared") or 0) for s in log)
    agg = {"checks": len(log), "scenarios": tot("B"), "decided": tot("decided"), "class_mismatch_decided": tot("class_mismatch_decided"),
           "class_mismatch_undecided": tot("class_mismatch_undecided"), "both_solved": tot("both_solved"),
           "compared_verified": tot("compared_verified"), "compared_iterate": tot("compared_iterate"), "exceptions": tot("exceptions"),
           "charged_to_oracle": tot("charged_to_oracle"), "iterate_outliers": tot("iterate_outliers"), "iters_compared": it_c,
           "iters_equal_frac": (it_eq / it_c) if it_c else None,
           "verified_max_err_in_tol_units": max([s.get("verified_max_err_in_tol_units") or 0.0 for s in log]),
           "iterate_max_err_in_tol_units": max([s.get("iterate_max_err_in_tol_units") or 0.0 for s in log])}
    tr.write_line(json.dumps({"parity_summary": agg}))
