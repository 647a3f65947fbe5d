import ctypes
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _c_double(L, *names):
    for n in names:
        getattr(L, n).restype = ctypes.c_double


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle port (oracle/hvapprox_oracle.c), built on demand."""
    path = os.path.join(ROOT, "oracle", "libhvoracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    L = ctypes.CDLL(path)
    _c_double(L, "orc_pow_uint", "orc_maxmin", "orc_expected_value", "orc_finalize", "orc_int_sin_pow",
              "orc_solve_inverse_u", "orc_hw_window", "orc_mc_window", "orc_rphi_window",
              "orc_hv_approx_hua_wang", "orc_hv_approx_normal", "orc_hv_approx_rphi_fang_wang_plus")
    L.orc_pow_uint.argtypes = [ctypes.c_double, ctypes.c_uint]
    L.orc_int_sin_pow.argtypes = [ctypes.c_int, ctypes.c_double]
    L.orc_solve_inverse_u.argtypes = [ctypes.c_double, ctypes.c_int]
    L.orc_finalize.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_uint64]
    L.orc_transform_and_filter.restype = ctypes.c_size_t
    return L


@pytest.fixture(scope="session")
def reflib():
    """The compiled reference (oracle/_ref/libhvref.so) when it is present; else skip."""
    path = os.path.join(ROOT, "oracle", "_ref", "libhvref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libhvref.so not built (needs /root/reference)")
    L = ctypes.CDLL(path)
    _c_double(L, "hv_approx_normal", "hv_approx_hua_wang", "hv_approx_rphi_fang_wang_plus", "ref_expected_value",
              "ref_hw_window", "ref_mc_window", "ref_rphi_window", "ref_pow_uint", "ref_int_sin_pow", "ref_solve_inverse_u")
    L.ref_pow_uint.argtypes = [ctypes.c_double, ctypes.c_uint]
    L.ref_int_sin_pow.argtypes = [ctypes.c_int, ctypes.c_double]
    L.ref_solve_inverse_u.argtypes = [ctypes.c_double, ctypes.c_int]
    L.ref_transform_and_filter.restype = ctypes.c_size_t
    return L


@pytest.fixture(scope="session")
def golden():
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    full = os.path.join(ROOT, "tests", "golden", "golden_full.json")
    if os.path.exists(full):
        F = json.load(open(full))
        for k in ("cfg3_full", "cfg3_full_windows"):
            if k in F:
                G[k] = F[k]
    return G


@pytest.fixture(scope="session")
def golden_r2():
    """Round-2 goldens (tests/golden/make_golden_r2.py): config 5 from the patched d <= 64 reference, DZ2019-HW at d = 21, 22."""
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_r2.json")))


@pytest.fixture(scope="session")
def mb():
    import moocore_b200

    return moocore_b200
