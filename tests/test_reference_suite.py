"""The reference's own tests of hv_approx, run against the oracle port (CPU) and the GPU product:
python/tests/test_moocore.py:532-578 (single point per dimension, default seed, error paths) and the dataset
known-answers of the doctest python/src/moocore/_moocore.py:1101-1130 (tests/golden/golden_datasets.json, written
by tests/golden/make_golden_datasets.py from the compiled reference)."""
import ctypes
import json
import os

import numpy as np
import pytest

from util import estimate, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATASETS = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_datasets.json")))
DOCTEST = {"input1": {"Rphi-FWE+": 93.5533, "DZ2019-HW": 93.5533}, "input1_filtered": {"Rphi-FWE+": 93.5533, "DZ2019-HW": 93.5533},
           "ran10pts9d_normalised": {"Rphi-FWE+": 0.48397532301, "DZ2019-HW": 0.483083547763}}


def dataset(name):
    rec = DATASETS[name]
    pts = np.array([[float.fromhex(v) for v in row] for row in rec["points"]])
    return pts, rec["ref"], rec["maximise"], rec["nsamples"], rec


@pytest.mark.parametrize("name", sorted(DATASETS))
@pytest.mark.parametrize("method", ["Rphi-FWE+", "DZ2019-HW"])
def test_oracle_dataset_known_answers(oracle, name, method):
    pts, ref, maxi, N, rec = dataset(name)
    got = estimate(oracle, "orc_", method, pts, ref, maxi, N)
    assert rel(got, float.fromhex(rec[method])) <= 1e-10
    assert abs(got - DOCTEST[name][method]) < 5e-5 * max(1.0, abs(got))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(DATASETS))
@pytest.mark.parametrize("method", ["Rphi-FWE+", "DZ2019-HW"])
def test_gpu_dataset_known_answers(mb, name, method):
    pts, ref, maxi, N, rec = dataset(name)
    got = mb.hv_approx(pts, ref=ref, maximise=maxi, nsamples=N, method=method)
    assert rel(got, float.fromhex(rec[method])) <= 1e-10
    # the doctest's own precision
    np.testing.assert_approx_equal(got, DOCTEST[name][method], significant=6 if name.startswith("input1") else 10)


@pytest.mark.gpu
@pytest.mark.parametrize("dim", range(2, 12))
def test_gpu_single_point_per_dimension(mb, dim):
    """python/tests/test_moocore.py:532-554 with the exact value 0.5**dim."""
    x = np.full((1, dim), 0.5)
    ref = np.full(dim, 1.0)
    true_hv = 0.5 ** dim
    np.testing.assert_approx_equal(true_hv, mb.hv_approx(x, ref=ref, method="DZ2019-MC", seed=42), significant=3 if dim < 7 else 2)
    signif = 4 if dim < 8 else 3 if dim < 10 else 2
    np.testing.assert_approx_equal(true_hv, mb.hv_approx(x, ref=ref, method="DZ2019-HW"), significant=signif)
    np.testing.assert_approx_equal(true_hv, mb.hv_approx(x, ref=ref), significant=signif)


@pytest.mark.gpu
def test_gpu_default_seed_and_maximise(mb):
    """python/tests/test_moocore.py:557-566"""
    x = np.full((1, 5), 0.5)
    ref = np.full(5, 0.0)
    appr = mb.hv_approx(x, ref=ref, method="DZ2019-MC", seed=None, maximise=True)
    np.testing.assert_approx_equal(0.5 ** 5, appr, significant=2)


@pytest.mark.gpu
def test_gpu_error_paths(mb):
    """python/tests/test_moocore.py:569-578, :143-144"""
    with pytest.raises(ValueError, match=r".*must be a positive integer value.*"):
        mb.hv_approx([[0, 0]], [1, 1], method="DZ2019-MC", nsamples="10")
    with pytest.raises(ValueError, match=r".*Unknown method.*"):
        mb.hv_approx([[0, 0]], [1, 1], method="None")
    assert mb.hv_approx([[1, 1, 1]], ref=1) == 0
    with pytest.raises(ValueError, match=r".*must have length 3.*"):
        mb.hv_approx([[0, 0, 0]], ref=[1, 1])
