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
