"""The reference's OWN callers running against libmoocore_b200.so.

dev/build_reference_binding.sh builds the reference's CFFI module with the INTEGRATION.md change applied
(python/src/moocore/_ffi_build.py: "hvapprox.c" out of `sources`, -lmoocore_b200 on the link line) and relinks the
command line tool c/main-hvapprox.c (c/Makefile:195) against the library; both are staged under
baseline/_ref/binding/ (git-ignored, shipped to the GPU box).  These tests then run, unmodified,
  * the reference's unit tests python/tests/test_moocore.py -k hv_approx (:532-578),
  * the hv_approx doctests of python/src/moocore/_moocore.py:1090-1130,
  * the three CLI known answers of SURVEY App. B,
against the swapped module.  Skipped when the staged build is absent (it needs /root/reference to be made)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "baseline", "_ref", "binding")


def _need_stage():
    if not os.path.exists(os.path.join(STAGE, "moocore", "_moocore.py")):
        pytest.skip("baseline/_ref/binding not built (run dev/build_reference_binding.sh where /root/reference exists)")


def _log(name, text):
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        open(os.path.join(out, name), "w").write(text)
    except OSError:
        pass


def _run(args, log):
    env = dict(os.environ, PYTHONPATH=STAGE)
    env.pop("HVB200_KERNEL", None)
    res = subprocess.run(args, cwd=STAGE, env=env, capture_output=True, text=True, timeout=900)
    _log(log, "$ " + " ".join(args) + "\n" + res.stdout + res.stderr)
    return res


def test_swapped_module_binds_the_library():
    _need_stage()
    code = ("import moocore, os\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libmoocore_b200.so' in maps, 'library not loaded'\n"
            "print(os.path.dirname(moocore.__file__))\n")
    res = _run([sys.executable, "-c", code], "r02_binding_import.log")
    assert res.returncode == 0, res.stdout + res.stderr
    assert os.path.realpath(res.stdout.strip().splitlines()[-1]) == os.path.realpath(os.path.join(STAGE, "moocore"))
    und = subprocess.run("nm -D --undefined-only moocore/_libmoocore*.so | grep hv_approx", shell=True, cwd=STAGE,
                         capture_output=True, text=True).stdout
    assert all(s in und for s in ("hv_approx_normal", "hv_approx_hua_wang", "hv_approx_rphi_fang_wang_plus"))


def test_reference_unit_tests_hv_approx():
    """python/tests/test_moocore.py:532-578, run as they are."""
    _need_stage()
    res = _run([sys.executable, "-m", "pytest", "tests/test_moocore.py", "-k", "hv_approx", "-q", "-rA"], "r02_binding_reference_unit_tests.log")
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "12 passed" in res.stdout


def test_reference_doctests_hv_approx():
    """python/src/moocore/_moocore.py:1090-1130 (doctest NUMBER flag as in python/pyproject.toml:210)."""
    _need_stage()
    res = _run([sys.executable, "-m", "pytest", "--doctest-modules", "moocore/_moocore.py", "-k", "hv_approx", "-q", "-rA"],
               "r02_binding_reference_doctests.log")
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "1 passed" in res.stdout


@pytest.mark.parametrize("m,extra,want", [(1, ["-S", "42"], "38.0008061008702"), (2, [], "37.9999580554839"), (3, [], "37.999978945303")])
def test_cli_known_answers(m, extra, want):
    """bin/hvapprox -q -r "10 10" -n 262144 -m {1,2,3} on the 4-point 2-D set, printed with %-22.15g."""
    _need_stage()
    res = _run([os.path.join(STAGE, "hvapprox_b200"), "-q", "-r", "10 10", "-n", "262144", "-m", str(m), *extra, "kat2d.dat"],
               f"r02_binding_cli_m{m}.log")
    assert res.returncode == 0, res.stdout + res.stderr
    got = res.stdout.strip().split()[0]
    assert abs(float(got) - float(want)) <= 1e-10 * float(want), (got, want)
    # the printed %-22.15g digits agree to 12+ significant digits: per-sample values are bit-exact, the sum is a tree
    # instead of the reference's left-to-right loop, so the 15th digit may differ
    assert got[:14] == want[:14] or abs(float(got) - float(want)) <= 2e-13 * float(want)


def test_cli_multiple_sets_and_no_dominating_point():
    """c/main-hvapprox.c:149-174 loops over the sets of a file; :167-168 turns a 0 return into a fatal error."""
    _need_stage()
    path = os.path.join(STAGE, "two_sets.dat")
    open(path, "w").write("5 5\n4 6\n\n2 7\n7 4\n")
    res = _run([os.path.join(STAGE, "hvapprox_b200"), "-q", "-r", "10 10", "-n", "10000", "-m", "2", "two_sets.dat"], "r02_binding_cli_sets.log")
    assert res.returncode == 0 and len(res.stdout.split()) == 2
    res = _run([os.path.join(STAGE, "hvapprox_b200"), "-q", "-r", "1 1", "-n", "1000", "-m", "2", "kat2d.dat"], "r02_binding_cli_zero.log")
    assert res.returncode != 0
