"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/*.h declares, and the Python surface validates arguments like the reference
(python/tests/test_moocore.py:143-144, 569-578).  No compute calls need a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for hdr in sorted(h for h in os.listdir(os.path.join(ROOT, "include")) if h.endswith(".h")):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)       # drop preprocessor lines (MOOCORE_API's own #define)
        for m in re.finditer(r"MOOCORE_API\s+[\w\s\*]+?\b(\w+)\s*\(", text):
            names.append(m.group(1))
    return names


def test_library_exports_all_declared_symbols(mb):
    names = declared_symbols()
    assert {"hv_approx_normal", "hv_approx_hua_wang", "hv_approx_rphi_fang_wang_plus", "epsilon_additive", "IGD",
            "avg_Hausdorff_dist", "whv_hype_unif", "whv_hype_expo", "whv_hype_gaus"} <= set(names)
    assert len(names) >= 15
    L = ctypes.CDLL(mb.library_path())
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported"


def test_abi_type_sizes():
    """dimension_t 1 byte, uint_fast32_t 8 bytes, boolvec 1 byte (non-R build)."""
    src = r'''
    #include "hvapprox.h"
    #include <stdio.h>
    int main(void){ printf("%zu %zu %zu %d\n", sizeof(dimension_t), sizeof(uint_fast32_t), sizeof(boolvec), MOOCORE_HVAPPROX_DIMENSION_MAX); return 0; }
    '''
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    assert out == ["1", "8", "1", "31"]


def test_version_and_device_count(mb):
    assert b"moocore_b200" in mb.lib.hvb200_version()
    assert mb.device_count() >= 0


def test_finalize_matches_oracle(mb, oracle):
    """hvb200_finalize is pure host arithmetic: c_d * sum / N in long double."""
    for d in (2, 3, 10, 20, 31):
        for s, N in ((123.456, 1000), (1e9, 100_000_000)):
            got = mb.hv_approx_finalize(d, N, [s, 0.0])
            assert got == oracle.orc_finalize(s, d, N)
    # ordered sum of several partials
    assert mb.hv_approx_finalize(3, 10, [1.0, 1e-17, 2.0, -1e-17]) == oracle.orc_finalize(3.0, 3, 10)


X = np.array([[5, 5], [4, 6], [2, 7], [7, 4]], float)


def test_python_surface_errors(mb):
    with pytest.raises(ValueError, match="must be a positive integer value"):
        mb.hv_approx(X, ref=[10, 10], nsamples=1000.5)
    with pytest.raises(ValueError, match="must be a positive integer value"):
        mb.hv_approx(X, ref=[10, 10], nsamples=0)
    with pytest.raises(ValueError, match="Unknown method"):
        mb.hv_approx(X, ref=[10, 10], method="does-not-exist")
    with pytest.raises(ValueError, match="must have length 2"):
        mb.hv_approx(X, ref=[10, 10, 10])
    with pytest.raises(ValueError, match="supports at most 31 columns"):
        mb.hv_approx(np.ones((3, 50)), ref=2.0)
    with pytest.raises(ValueError, match="at least 1 column"):
        mb.hv_approx(np.ones((3, 0)), ref=2.0)


def test_one_objective_is_exact(mb):
    """nobj == 1 returns the exact hypervolume (python/src/moocore/_moocore.py:1139-1141)."""
    assert mb.hv_approx(np.array([[3.0], [1.0], [2.0]]), ref=5.0) == 4.0
    assert mb.hv_approx(np.array([[3.0], [1.0]]), ref=0.0) == 0.0
    assert mb.hv_approx(np.array([[3.0], [1.0]]), ref=0.0, maximise=True) == 3.0


def test_no_gpu_fails_loudly(mb):
    if mb.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mb.hv_approx(X, ref=[10, 10], method="DZ2019-HW")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mb.Plan(X, [10, 10])


def test_product_does_not_reference_the_oracle():
    """The product path may not include, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "moocore_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for bad in ("libhvoracle", "libhvref", "orc_", "oracle_tables.h", "/root/reference\""):
                    assert bad not in text, f"{f} mentions {bad}"
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libmoocore_b200.so")], capture_output=True, text=True).stdout
    assert "hvoracle" not in out and "hvref" not in out


@pytest.mark.parametrize("d", [2, 3, 4, 5, 8, 10, 16, 20, 31, 32])
def test_rphi_integer_recurrence_matches_x87(mb, oracle, d):
    """hvb200_rphi_sequence (the library's threaded-per-objective host recurrence, objective-major
    output) against the oracle's sequential recurrence: bit-exact, deep into the sequence."""
    import ctypes
    from util import P
    L = mb.lib
    L.hvb200_rphi_sequence.restype = ctypes.c_int
    L.hvb200_rphi_sequence.argtypes = [ctypes.c_uint8, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double)]
    for i0, cnt in ((0, 5000), (300_000, 40_000)):
        got = np.zeros((cnt, d - 1))
        assert L.hvb200_rphi_sequence(d, i0, cnt, P(got)) == 0
        want = np.zeros((cnt, d - 1))
        oracle.orc_rphi_window(None, ctypes.c_size_t(0), d, ctypes.c_uint64(i0), ctypes.c_uint64(cnt), None, None, P(want))
        assert np.array_equal(got, want), (d, i0)
