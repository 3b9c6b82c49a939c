"""Shared test plumbing.

Markers: `gpu` tests need a real CUDA device and call the product library
(mappy-rs_b200/libmmg.so) through its C ABI; everything else runs on CPU:
the oracle against the reference's fixtures, the host logic, the ABI surface,
and the kernel SOURCE executed lane-for-lane by the test-only SIMT emulator
(tests/emu), which is never loaded by the product.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mappy-rs_b200")
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
PRODUCT_LIB = os.path.join(PKG, "libmmg.so")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libmmg_emu.so")


# the reference's own test file is kept verbatim under tests/golden/reference_tests and is run, unchanged, by
# tests/test_reference_pytests.py in a separate pytest process against mappy-rs_b200/mappy_rs (GPU box only)
collect_ignore_glob = ["golden/*"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j8"])
    import mm2oracle
    return mm2oracle


@pytest.fixture(scope="session")
def emu_lib():
    """The product's kernel sources compiled by g++ against the SIMT emulator (test build)."""
    if os.path.isdir(os.path.join(PKG, "csrc")) and subprocess.call(["which", "g++"], stdout=subprocess.DEVNULL) == 0:
        subprocess.check_call(["make", "-s", "-C", PKG, "emu", "-j8"])
    from mappy_rs import _mmg
    return _mmg.Lib(EMU_LIB)


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library (nvcc, sm_100a). No fallback: missing library = error."""
    from mappy_rs import _mmg
    return _mmg.Lib(PRODUCT_LIB)
