"""moocore_b200 — B200-native drop-in for moocore's ``hv_approx`` hot path.

Host-side mirror of ``moocore.hv_approx`` (python/src/moocore/_moocore.py:991-1177): same name,
argument meaning, defaults, validation and error messages; the call lands in the same three C
symbols (c/hvapprox.h:16-38), here served by ``libmoocore_b200.so`` (hand-written sm_100a CUDA
kernels behind a C-ABI, see include/hvapprox.h).  There is no CPU fallback: importing works
without a GPU, computing raises.
"""
from ._api import (  # noqa: F401
    HVAPPROX_DIMENSION_MAX,
    hv_approx,
    hv_approx_range,
    hv_approx_sets,
    hv_contributions_approx,
    hv_approx_shard,
    hv_approx_finalize,
    Plan,
    lib,
    library_path,
    device_count,
    directions,
    epsilon_additive,
    epsilon_mult,
    igd,
    igd_plus,
    avg_hausdorff_dist,
    whv_hype,
    measure_fp64,
)

__all__ = [
    "hv_approx",
    "hv_approx_range",
    "hv_approx_sets",
    "hv_contributions_approx",
    "hv_approx_shard",
    "hv_approx_finalize",
    "Plan",
    "HVAPPROX_DIMENSION_MAX",
    "lib",
    "library_path",
    "device_count",
    "directions",
    "epsilon_additive",
    "epsilon_mult",
    "igd",
    "igd_plus",
    "avg_hausdorff_dist",
    "whv_hype",
    "measure_fp64",
]
__version__ = "0.1.0"
