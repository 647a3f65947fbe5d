"""ctypes binding of libmoocore_b200.so and the Python surface on top of it."""
from __future__ import annotations

import ctypes
import os
from typing import Any, Literal, Sequence

import numpy as np

HVAPPROX_DIMENSION_MAX = 31  # MOOCORE_HVAPPROX_DIMENSION_MAX, c/config.h:28

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBNAME = "libmoocore_b200.so"

METHOD_IDS = {"DZ2019-MC": 1, "DZ2019-HW": 2, "Rphi-FWE+": 3}
KERNEL_IDS = {"auto": 0, "exact": 1, "screen": 2, "dpx": 3, "pruned": 4}

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_u8_p = ctypes.POINTER(ctypes.c_uint8)


class Stats(ctypes.Structure):
    """struct hvb200_stats (include/hvapprox_b200.h)."""

    _fields_ = [
        ("maxmin_ms", ctypes.c_double),
        ("dirgen_ms", ctypes.c_double),
        ("reduce_ms", ctypes.c_double),
        ("h2d_ms", ctypes.c_double),
        ("maxmin_launches", ctypes.c_uint64),
        ("dirgen_launches", ctypes.c_uint64),
        ("reduce_launches", ctypes.c_uint64),
        ("samples", ctypes.c_uint64),
        ("npoints", ctypes.c_uint64),
        ("newton_capped", ctypes.c_uint64),
        ("slow_path", ctypes.c_uint64),
        ("h2d_bytes", ctypes.c_uint64),
        ("kernel", ctypes.c_int),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


class McSegment(ctypes.Structure):
    """hvb200_mc_segment of include/hvapprox_b200.h: what a rank knows about its pair range of the MC stream."""

    _fields_ = [
        ("pair0", ctypes.c_uint64),
        ("npairs", ctypes.c_uint64),
        ("merge", ctypes.c_uint64),
        ("head", ctypes.c_uint64 * 16),
        ("body", ctypes.c_uint64),
        ("skip_out", ctypes.c_uint64),
    ]
    WORDS = 21

    def to_list(self) -> list:
        return [self.pair0, self.npairs, self.merge, *self.head, self.body, self.skip_out]

    @classmethod
    def from_list(cls, v) -> "McSegment":
        v = [int(x) for x in v]
        return cls(v[0], v[1], v[2], (ctypes.c_uint64 * 16)(*v[3:19]), v[19], v[20])


class McStart(ctypes.Structure):
    """hvb200_mc_start: where a rank's parse starts, what it drops, and its sample range."""

    _fields_ = [("pair", ctypes.c_uint64), ("discard", ctypes.c_uint64), ("i0", ctypes.c_uint64), ("i1", ctypes.c_uint64)]


MC_OPEN_END = (1 << 64) - 1


def library_path() -> str:
    # HVB200_LIBRARY: development override (A/B-testing differently tuned builds)
    return os.environ.get("HVB200_LIBRARY") or os.path.join(_HERE, _LIBNAME)


_lib = None


def _load() -> ctypes.CDLL:
    """Load the native library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C moocore_b200/csrc`. moocore_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(path)
    dim_t, size_t, u64, u32 = ctypes.c_uint8, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32
    vp = ctypes.c_void_p
    common = [_c_double_p, size_t, dim_t, _c_double_p, _c_u8_p, u64]
    L.hv_approx_normal.restype = ctypes.c_double
    L.hv_approx_normal.argtypes = common + [u32]
    L.hv_approx_hua_wang.restype = ctypes.c_double
    L.hv_approx_hua_wang.argtypes = common
    L.hv_approx_rphi_fang_wang_plus.restype = ctypes.c_double
    L.hv_approx_rphi_fang_wang_plus.argtypes = common
    L.hvb200_last_error.restype = ctypes.c_char_p
    L.hvb200_version.restype = ctypes.c_char_p
    L.hvb200_device_count.restype = ctypes.c_int
    L.hvb200_plan_create.restype = ctypes.c_int
    L.hvb200_plan_create.argtypes = [ctypes.POINTER(vp), _c_double_p, size_t, dim_t, _c_double_p, _c_u8_p, ctypes.c_int]
    L.hvb200_plan_destroy.restype = None
    L.hvb200_plan_destroy.argtypes = [vp]
    L.hvb200_plan_npoints.restype = size_t
    L.hvb200_plan_npoints.argtypes = [vp]
    L.hvb200_plan_ndominated.restype = size_t
    L.hvb200_plan_ndominated.argtypes = [vp]
    L.hvb200_plan_set_kernel.restype = ctypes.c_int
    L.hvb200_plan_set_kernel.argtypes = [vp, ctypes.c_int]
    L.hvb200_plan_run.restype = ctypes.c_int
    L.hvb200_plan_run.argtypes = [vp, ctypes.c_int, u64, u32, u64, u64, vp, _c_double_p]
    for nm in ("epsilon_additive", "epsilon_mult"):
        getattr(L, nm).restype = ctypes.c_double
        getattr(L, nm).argtypes = [_c_double_p, ctypes.c_size_t, ctypes.c_uint8, _c_double_p, ctypes.c_size_t, _c_u8_p]
    for nm in ("IGD", "IGD_plus"):
        getattr(L, nm).restype = ctypes.c_double
        getattr(L, nm).argtypes = [_c_double_p, ctypes.c_size_t, ctypes.c_uint8, _c_double_p, ctypes.c_size_t, _c_u8_p]
    L.avg_Hausdorff_dist.restype = ctypes.c_double
    L.avg_Hausdorff_dist.argtypes = [_c_double_p, ctypes.c_size_t, ctypes.c_uint8, _c_double_p, ctypes.c_size_t, _c_u8_p, ctypes.c_uint]
    for nm, extra in (("whv_hype_unif", []), ("whv_hype_expo", [ctypes.c_double]), ("whv_hype_gaus", [_c_double_p])):
        getattr(L, nm).restype = ctypes.c_double
        getattr(L, nm).argtypes = [_c_double_p, ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, u32] + extra
    L.hvb200_whv_ordered_sum.restype = ctypes.c_double
    L.hvb200_whv_ordered_sum.argtypes = [ctypes.POINTER(ctypes.c_uint), ctypes.c_size_t, ctypes.c_int]
    L.hvb200_hv_approx_sets.restype = ctypes.c_int
    L.hvb200_hv_approx_sets.argtypes = [ctypes.c_int, _c_double_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t, ctypes.c_uint8,
                                        _c_double_p, _c_u8_p, u64, u32, ctypes.c_int, _c_double_p]
    L.hvb200_hv_contributions_approx.restype = ctypes.c_int
    L.hvb200_hv_contributions_approx.argtypes = [ctypes.c_int, _c_double_p, size_t, dim_t, _c_double_p, _c_u8_p, u64, u32, ctypes.c_int,
                                                 ctypes.c_int, _c_double_p, ctypes.POINTER(ctypes.c_double)]
    L.hvb200_multi_last_exchange.restype = ctypes.c_int
    L.hvb200_plan_run_blocks.restype = ctypes.c_int
    L.hvb200_plan_run_blocks.argtypes = [vp, ctypes.c_int, u64, u32, u64, u64, u64, vp, _c_double_p]
    L.hvb200_mc_segment_layout.restype = ctypes.c_int
    L.hvb200_mc_segment_layout.argtypes = [u64, dim_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.hvb200_plan_mc_segment_scan.restype = ctypes.c_int
    L.hvb200_plan_mc_segment_scan.argtypes = [vp, u32, u64, u64, ctypes.POINTER(McSegment)]
    L.hvb200_mc_segments_resolve.restype = ctypes.c_int
    L.hvb200_mc_segments_resolve.argtypes = [ctypes.POINTER(McSegment), ctypes.c_int, dim_t, u64, ctypes.POINTER(McStart)]
    L.hvb200_plan_set_mc_start.restype = ctypes.c_int
    L.hvb200_plan_set_mc_start.argtypes = [vp, u64, u64]
    L.hvb200_plan_last_stats.restype = ctypes.c_int
    L.hvb200_plan_last_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.hvb200_finalize.restype = ctypes.c_double
    L.hvb200_finalize.argtypes = [dim_t, u64, _c_double_p, ctypes.c_int]
    L.hvb200_plan_sample_values.restype = ctypes.c_int
    L.hvb200_plan_sample_values.argtypes = [vp, ctypes.c_int, u64, u32, u64, u64, _c_double_p]
    L.hvb200_directions.restype = ctypes.c_int
    L.hvb200_directions.argtypes = [ctypes.c_int, dim_t, u64, u32, u64, u64, _c_double_p, ctypes.c_int]
    L.hvb200_plan_run_directions.restype = ctypes.c_int
    L.hvb200_plan_run_directions.argtypes = [vp, _c_double_p, u64, _c_double_p, _c_double_p]
    L.hvb200_hv_approx_multi.restype = ctypes.c_double
    L.hvb200_hv_approx_multi.argtypes = [ctypes.c_int, _c_double_p, size_t, dim_t, _c_double_p, _c_u8_p, u64, u32, ctypes.c_int]
    L.hvb200_measure_fp64_flops.restype = ctypes.c_double
    L.hvb200_measure_fp64_flops.argtypes = [ctypes.c_int, _c_double_p, _c_double_p]
    _lib = L
    return L


class _LazyLib:
    def __getattr__(self, name):
        return getattr(_load(), name)


lib = _LazyLib()


def device_count() -> int:
    return int(_load().hvb200_device_count())


def _last_error() -> str:
    return (_load().hvb200_last_error() or b"").decode("utf-8", "replace")


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_c_double_p)


# ---- argument handling: same rules and messages as python/src/moocore/_utils.py ------------
def _array_1d_of_length_n(x: Any, n: int, name: str = "x") -> np.ndarray:
    """_utils.py:64-72."""
    x = np.ravel(x)
    if len(x) == 1:
        return np.full((n), x[0])
    if x.shape[0] == n:
        return x
    raise ValueError(f"{name!r} must have length {n}, but it has length {x.shape[0]}")


def _is_integer_value(n: Any) -> bool:
    """_utils.py:75-86."""
    if isinstance(n, (int, np.integer)):
        return True
    if n is None:
        return False
    try:
        return int(n) == n
    except (ValueError, TypeError):
        return False


def _get_seed_for_c(seed: Any) -> int:
    """_utils.py:89-92: int -> as is; None / Generator -> draw from default_rng."""
    if not _is_integer_value(seed):
        seed = np.random.default_rng(seed).integers(2**32 - 2, dtype=np.uint32)
    return int(seed)


def _prepare(points, ref, maximise, nsamples, allow_extension=False):
    points = np.asarray(points, dtype=float)
    if points.ndim != 2:
        raise ValueError("input points must be a 2-D array")
    nobj = points.shape[1]
    if nobj == 0:
        raise ValueError("input points must have at least 1 column")
    max_dim = 64 if allow_extension else HVAPPROX_DIMENSION_MAX
    if nobj > max_dim:
        raise ValueError(f"This function supports at most {max_dim} columns, but input has {nobj}")
    ref = np.ascontiguousarray(_array_1d_of_length_n(np.asarray(ref, dtype=float), nobj, name="ref"), dtype=float)
    if not _is_integer_value(nsamples) or nsamples <= 0 or nsamples > 2147483648:
        raise ValueError(f"nsamples ({nsamples}) must be a positive integer value smaller than 2147483648")
    maxi = np.ascontiguousarray(_array_1d_of_length_n(maximise, nobj, name="maximise").astype(bool)).view(np.uint8)
    points = np.ascontiguousarray(points)
    return points, ref, maxi, int(nsamples), nobj


def _exact_hv_1d(points: np.ndarray, ref: np.ndarray, maximise: np.ndarray) -> float:
    """nobj == 1: the reference returns hypervolume() (_moocore.py:1139-1141); in 1-D that is the
    distance from ref to the best point (0 if none dominates ref)."""
    x = points[:, 0]
    if x.size == 0:
        return 0.0
    v = (x.max() - ref[0]) if maximise[0] else (ref[0] - x.min())
    return float(max(v, 0.0))


def hv_approx(
    points,
    /,
    ref,
    *,
    maximise: bool | Sequence[bool] = False,
    nsamples: int = 262_144,
    seed: int | np.random.Generator | None = None,
    method: Literal["DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"] = "Rphi-FWE+",
) -> float:
    """Approximate the hypervolume indicator (drop-in for ``moocore.hv_approx``).

    Same contract as python/src/moocore/_moocore.py:991-1177: ``points`` is (n, m) float,
    ``ref`` a scalar or length-m vector, ``maximise`` a bool or length-m bools, ``nsamples`` a
    positive integer value <= 2**31, ``seed`` only used by ``"DZ2019-MC"``.  Returns ``0.0`` when
    no point strictly dominates ``ref``.  Raises ``ValueError`` for a non-integer ``nsamples``,
    an unknown ``method``, a wrong-length ``ref`` or more than 31 columns, like the reference.
    Raises ``RuntimeError`` if the CUDA path fails (the C entry point returned NaN).
    """
    if method not in METHOD_IDS:
        # validated up front; the reference raises the same error from its match statement
        pts = np.asarray(points, dtype=float)
        if pts.ndim == 2 and pts.shape[1] >= 2:
            _prepare(points, ref, maximise, nsamples)
        raise ValueError(f"Unknown method = {method}")
    points, ref, maxi, nsamples, nobj = _prepare(points, ref, maximise, nsamples)
    if nobj == 1:
        return _exact_hv_1d(points, ref, maxi)
    L = _load()
    args = (_dp(points), points.shape[0], nobj, _dp(ref), maxi.ctypes.data_as(_c_u8_p), nsamples)
    if method == "DZ2019-MC":
        hv = L.hv_approx_normal(*args, _get_seed_for_c(seed))
    elif method == "DZ2019-HW":
        hv = L.hv_approx_hua_wang(*args)
    else:
        hv = L.hv_approx_rphi_fang_wang_plus(*args)
    if hv != hv:
        raise RuntimeError(f"moocore_b200: CUDA hv_approx failed: {_last_error()}")
    return float(hv)


def hv_approx_sets(
    sets,
    /,
    ref,
    *,
    maximise: bool | Sequence[bool] = False,
    nsamples: int = 262_144,
    seed: int | np.random.Generator | None = None,
    method: Literal["DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"] = "Rphi-FWE+",
    device: int = 0,
) -> np.ndarray:
    """``hv_approx`` of many point sets that share ``ref``, ``maximise`` and ONE direction set, in one call.

    ``sets`` is a sequence of (n_i, m) arrays (or a ``(data, cumsizes)`` pair in the layout of the reference's
    command line tool, c/main-hvapprox.c:149-174).  Returns one estimate per set, each equal to
    ``hv_approx(set_i, ...)`` up to the summation order; 0.0 for sets without a point that strictly dominates
    ``ref``.  The directions are generated once and small sets share one kernel launch."""
    if method not in METHOD_IDS:
        raise ValueError(f"Unknown method = {method}")
    if isinstance(sets, tuple) and len(sets) == 2 and np.ndim(sets[1]) == 1 and np.ndim(sets[0]) == 2:
        data = np.ascontiguousarray(sets[0], dtype=float)
        cum = np.ascontiguousarray(sets[1], dtype=np.uint64)
        if cum.size and (np.any(np.diff(cum.astype(np.int64)) < 0) or int(cum[-1]) > data.shape[0]):
            raise ValueError("cumsizes must be non-decreasing and its last entry at most the number of rows of data")
    else:
        arrs = [np.ascontiguousarray(a, dtype=float) for a in sets]
        if not arrs:
            return np.zeros(0)
        if any(a.ndim != 2 or a.shape[1] != arrs[0].shape[1] for a in arrs):
            raise ValueError("all sets must be 2-D arrays with the same number of columns")
        data = np.ascontiguousarray(np.concatenate(arrs, axis=0))
        cum = np.cumsum([a.shape[0] for a in arrs]).astype(np.uint64)
    probe = data if data.shape[0] else np.zeros((1, data.shape[1]))
    _, ref, maxi, nsamples, nobj = _prepare(probe, ref, maximise, nsamples)
    if nobj < 2:
        raise ValueError("hv_approx_sets needs at least 2 objectives")
    out = np.zeros(cum.size)
    L = _load()
    rc = L.hvb200_hv_approx_sets(METHOD_IDS[method], _dp(data), cum.ctypes.data_as(ctypes.POINTER(ctypes.c_size_t)), cum.size,
                                 nobj, _dp(ref), maxi.ctypes.data_as(_c_u8_p), nsamples,
                                 _get_seed_for_c(seed) if method == "DZ2019-MC" else 0, device, _dp(out))
    if rc != 0:
        raise RuntimeError(f"moocore_b200: hv_approx_sets failed ({rc}): {_last_error()}")
    return out


def hv_contributions_approx(
    points,
    /,
    ref,
    *,
    maximise: bool | Sequence[bool] = False,
    nsamples: int = 262_144,
    seed: int | np.random.Generator | None = None,
    method: Literal["DZ2019-HW", "DZ2019-MC", "Rphi-FWE+"] = "Rphi-FWE+",
    ignore_dominated: bool = True,
    device: int = 0,
    return_hv: bool = False,
):
    """Approximate hypervolume contributions ``HV(P) - HV(P \\ {p_i})`` of every row of ``points`` from the direction
    set of ``hv_approx`` (same arguments).  For each direction the point attaining ``max_p min_k`` is credited with
    ``s1^d - s2^d`` (largest and second largest value), so the contributions and the hypervolume estimate come from ONE
    pass.  Rows that do not strictly dominate ``ref`` (and duplicates) get 0.  ``ignore_dominated`` as in
    ``moocore.hv_contributions`` (_moocore.py): True removes strictly dominated rows first (they get 0 and do not shield
    the rows that dominate them), False keeps them in the set.  Extension: the reference only computes
    exact contributions (``moocore.hv_contributions``); with ``return_hv=True`` the estimate of the hypervolume itself
    is returned as well."""
    if method not in METHOD_IDS:
        raise ValueError(f"Unknown method = {method}")
    points, ref, maxi, nsamples, nobj = _prepare(points, ref, maximise, nsamples)
    if nobj < 2:
        raise ValueError("hv_contributions_approx needs at least 2 objectives")
    out = np.zeros(points.shape[0])
    hv = ctypes.c_double(0.0)
    L = _load()
    rc = L.hvb200_hv_contributions_approx(METHOD_IDS[method], _dp(points), points.shape[0], nobj, _dp(ref), maxi.ctypes.data_as(_c_u8_p),
                                          nsamples, _get_seed_for_c(seed) if method == "DZ2019-MC" else 0, device, 1 if ignore_dominated else 0,
                                          _dp(out), ctypes.byref(hv))
    if rc != 0:
        raise RuntimeError(f"moocore_b200: hv_contributions_approx failed ({rc}): {_last_error()}")
    return (out, float(hv.value)) if return_hv else out


def _refset_common(points, ref, maximise, check_all_positive=False):
    """Argument handling of python/src/moocore/_moocore.py:275-310 (_unary_refset_common)."""
    points = np.ascontiguousarray(np.asarray(points, dtype=float))
    ref = np.ascontiguousarray(np.atleast_2d(np.asarray(ref, dtype=float)))
    if points.ndim != 2:
        raise ValueError("input points must be a 2-D array")
    if check_all_positive:
        if not (points > 0).all():
            raise ValueError("All values must be larger than 0 in the input points")
        if not (ref > 0).all():
            raise ValueError("All values must be larger than 0 in the reference set")
    nobj = points.shape[1]
    if nobj != ref.shape[1]:
        raise ValueError(f"points and ref need to have the same number of columns ({nobj} != {ref.shape[1]})")
    if nobj == 0:
        raise ValueError("The number of columns cannot be 0")
    maxi = np.ascontiguousarray(np.broadcast_to(np.asarray(maximise, dtype=bool), (nobj,))).view(np.uint8)
    return points, ref, maxi, nobj


def _epsilon(name, points, ref, maximise, check_all_positive):
    points, ref, maxi, nobj = _refset_common(points, ref, maximise, check_all_positive)
    if points.shape[0] == 0 or ref.shape[0] == 0:
        raise ValueError("points and ref must not be empty")
    if nobj > 64:
        raise ValueError("This function supports at most 64 columns")
    fn = getattr(_load(), name)
    v = fn(_dp(points), points.shape[0], nobj, _dp(ref), ref.shape[0], maxi.ctypes.data_as(_c_u8_p))
    if v != v:
        raise RuntimeError(f"moocore_b200: CUDA {name} failed: {_last_error()}")
    return float(v)


def epsilon_additive(points, /, ref, *, maximise: bool | Sequence[bool] = False) -> float:
    """Additive epsilon indicator (drop-in for ``moocore.epsilon_additive``, _moocore.py:462-511)."""
    return _epsilon("epsilon_additive", points, ref, maximise, False)


def epsilon_mult(points, /, ref, *, maximise: bool | Sequence[bool] = False) -> float:
    """Multiplicative epsilon indicator (drop-in for ``moocore.epsilon_mult``); all values must be > 0."""
    return _epsilon("epsilon_mult", points, ref, maximise, True)


def _gd_call(name, points, ref, maximise, *extra):
    points, ref, maxi, nobj = _refset_common(points, ref, maximise)
    if ref.shape[0] == 0:
        raise ValueError("ref must not be empty")
    if nobj > 64:
        raise ValueError("This function supports at most 64 columns")
    v = getattr(_load(), name)(_dp(points), points.shape[0], nobj, _dp(ref), ref.shape[0], maxi.ctypes.data_as(_c_u8_p), *extra)
    if v != v:
        raise RuntimeError(f"moocore_b200: CUDA {name} failed: {_last_error()}")
    return float(v)


def igd(points, /, ref, *, maximise: bool | Sequence[bool] = False) -> float:
    """Inverted generational distance (drop-in for ``moocore.igd``, _moocore.py:310-390)."""
    return _gd_call("IGD", points, ref, maximise)


def igd_plus(points, /, ref, *, maximise: bool | Sequence[bool] = False) -> float:
    """Modified inverted generational distance IGD+ (drop-in for ``moocore.igd_plus``)."""
    return _gd_call("IGD_plus", points, ref, maximise)


def avg_hausdorff_dist(points, /, ref, *, maximise: bool | Sequence[bool] = False, p: int = 1) -> float:
    """Averaged Hausdorff distance (drop-in for ``moocore.avg_hausdorff_dist``, _moocore.py:416-460)."""
    if p <= 0:
        raise ValueError("'p' must be larger than zero")
    return _gd_call("avg_Hausdorff_dist", points, ref, maximise, ctypes.c_uint(p))


def whv_hype(points, /, *, ref, ideal, maximise: bool | Sequence[bool] = False, nsamples: int = 100000,
             seed: int | np.random.Generator | None = None, dist: str = "uniform", mu=None) -> float:
    """HypE Monte-Carlo estimate of the (weighted) hypervolume, 2 objectives (drop-in for ``moocore.whv_hype``,
    _moocore.py:2767-2912): ``dist`` is ``"uniform"``, ``"exponential"`` (``mu`` a float) or ``"point"`` (``mu`` a
    goal point).  Bit-identical to the reference for the same seed."""
    points = np.array(np.asarray(points, dtype=float), order="C", copy=True)
    if points.ndim != 2:
        raise ValueError("input points must be a 2-D array")
    nobj = points.shape[1]
    if nobj != 2:
        raise NotImplementedError("Only 2D points are currently supported")
    ref = np.array(_array_1d_of_length_n(np.asarray(ref, dtype=float), nobj, name="ref"), dtype=float)
    ideal = np.array(_array_1d_of_length_n(np.asarray(ideal, dtype=float), nobj, name="ideal"), dtype=float)
    maxi = np.broadcast_to(np.asarray(maximise, dtype=bool), (nobj,))
    if maxi.any():                                  # _moocore.py:2878-2886: negate the maximised objectives
        points[:, maxi] = -points[:, maxi]
        ref[maxi] = -ref[maxi]
        ideal[maxi] = -ideal[maxi]
    L = _load()
    args = (_dp(points), points.shape[0], _dp(ideal), _dp(ref), int(nsamples), _get_seed_for_c(seed))
    if dist == "uniform":
        v = L.whv_hype_unif(*args)
    elif dist == "exponential":
        v = L.whv_hype_expo(*args, float(mu))
    elif dist == "point":
        mu = np.array(_array_1d_of_length_n(np.asarray(mu, dtype=float), nobj, name="mu"), dtype=float)
        v = L.whv_hype_gaus(*args, _dp(mu))
    else:
        raise ValueError(f"Unknown value of dist = {dist}")
    if v != v and _last_error():
        raise RuntimeError(f"moocore_b200: CUDA whv_hype failed: {_last_error()}")
    return float(v)


class Plan:
    """A point set transformed, filtered and resident on one GPU (hvb200_plan)."""

    def __init__(self, points, ref, maximise=False, device: int = 0, kernel: str = "auto", allow_extension: bool = False):
        points, ref, maxi, _, nobj = _prepare(points, ref, maximise, 1, allow_extension)
        if nobj < 2:
            raise ValueError("Plan needs at least 2 objectives")
        L = _load()
        h = ctypes.c_void_p()
        rc = L.hvb200_plan_create(ctypes.byref(h), _dp(points), points.shape[0], nobj, _dp(ref),
                                  maxi.ctypes.data_as(_c_u8_p), device)
        if rc != 0:
            raise RuntimeError(f"hvb200_plan_create failed ({rc}): {_last_error()}")
        self._h = h if h.value else None
        self.nobj = nobj
        self.device = device
        if self._h is not None and kernel != "auto":
            self.set_kernel(kernel)

    @property
    def empty(self) -> bool:
        """True when no point strictly dominates ref (hv_approx is then 0)."""
        return self._h is None

    @property
    def npoints(self) -> int:
        return 0 if self._h is None else int(_load().hvb200_plan_npoints(self._h))

    @property
    def ndominated(self) -> int:
        """Points the nondominated pre-filter removed at plan creation (they cannot change any sample value)."""
        return 0 if self._h is None else int(_load().hvb200_plan_ndominated(self._h))

    def set_kernel(self, kernel: str) -> None:
        if kernel not in KERNEL_IDS:
            raise ValueError(f"Unknown kernel = {kernel}")
        if self._h is None:
            return
        if _load().hvb200_plan_set_kernel(self._h, KERNEL_IDS[kernel]) != 0:
            raise RuntimeError(_last_error())

    def run(self, method: str, nsamples: int, seed: int = 0, i0: int = 0, i1: int | None = None, stream: int | None = None):
        """Partial sum over samples [i0, i1) as a (hi, lo) pair."""
        if self._h is None:
            return (0.0, 0.0)
        i1 = nsamples if i1 is None else i1
        out = (ctypes.c_double * 2)()
        rc = _load().hvb200_plan_run(self._h, METHOD_IDS[method], nsamples, seed, i0, i1,
                                     ctypes.c_void_p(stream) if stream else None, out)
        if rc != 0:
            raise RuntimeError(f"hvb200_plan_run failed ({rc}): {_last_error()}")
        return (out[0], out[1])

    def run_blocks(self, method: str, nsamples: int, first: int, stride: int, block: int = 1 << 14, seed: int = 0,
                   stream: int | None = None):
        """Partial sum over the block-cyclic shard `first` of `stride` (blocks of `block` directions; DZ2019-HW)."""
        if self._h is None:
            return (0.0, 0.0)
        out = (ctypes.c_double * 2)()
        rc = _load().hvb200_plan_run_blocks(self._h, METHOD_IDS[method], nsamples, seed, block, first, stride,
                                            ctypes.c_void_p(stream) if stream else None, out)
        if rc != 0:
            raise RuntimeError(f"hvb200_plan_run_blocks failed ({rc}): {_last_error()}")
        return (out[0], out[1])

    def mc_segment_scan(self, seed: int, pair0: int, npairs: int) -> McSegment:
        """Describe the pair range [pair0, pair0 + npairs) of the DZ2019-MC stream (hvb200_plan_mc_segment_scan);
        npairs = MC_OPEN_END for the last rank's open-ended range."""
        if self._h is None:
            raise RuntimeError("empty plan")
        seg = McSegment()
        rc = _load().hvb200_plan_mc_segment_scan(self._h, seed, pair0, npairs, ctypes.byref(seg))
        if rc != 0:
            raise RuntimeError(f"hvb200_plan_mc_segment_scan failed ({rc}): {_last_error()}")
        return seg

    def set_mc_start(self, pair: int, discard: int) -> None:
        """The next DZ2019-MC run of this plan starts its parse at `pair` and drops `discard` normals first."""
        if self._h is not None and _load().hvb200_plan_set_mc_start(self._h, pair, discard) != 0:
            raise RuntimeError(_last_error())

    def stats(self) -> dict:
        s = Stats()
        if self._h is not None:
            _load().hvb200_plan_last_stats(self._h, ctypes.byref(s))
        return s.as_dict()

    def sample_values(self, method: str, nsamples: int, seed: int = 0, i0: int = 0, count: int | None = None) -> np.ndarray:
        count = nsamples - i0 if count is None else count
        out = np.zeros(count, dtype=np.float64)
        if self._h is None:
            return out
        rc = _load().hvb200_plan_sample_values(self._h, METHOD_IDS[method], nsamples, seed, i0, count, _dp(out))
        if rc != 0:
            raise RuntimeError(f"hvb200_plan_sample_values failed ({rc}): {_last_error()}")
        return out

    def run_directions(self, r: np.ndarray, want_values: bool = True):
        r = np.ascontiguousarray(r, dtype=np.float64)
        assert r.ndim == 2 and r.shape[1] == self.nobj
        vals = np.zeros(r.shape[0], dtype=np.float64) if want_values else None
        if self._h is None:
            return vals, (0.0, 0.0)
        out = (ctypes.c_double * 2)()
        rc = _load().hvb200_plan_run_directions(self._h, _dp(r), r.shape[0], _dp(vals) if want_values else None, out)
        if rc != 0:
            raise RuntimeError(f"hvb200_plan_run_directions failed ({rc}): {_last_error()}")
        return vals, (out[0], out[1])

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            _load().hvb200_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def directions(method: str, nobj: int, nsamples: int, seed: int = 0, i0: int = 0, count: int | None = None, device: int = 0) -> np.ndarray:
    """Reciprocal directions r^(i) generated on the device (parity/debug)."""
    count = nsamples - i0 if count is None else count
    out = np.zeros((count, nobj), dtype=np.float64)
    rc = _load().hvb200_directions(METHOD_IDS[method], nobj, nsamples, seed, i0, count, _dp(out), device)
    if rc != 0:
        raise RuntimeError(f"hvb200_directions failed ({rc}): {_last_error()}")
    return out


def mc_segment_layout(nsamples: int, nobj: int, world: int, rank: int):
    """(pair0, npairs) of rank's pair range of the DZ2019-MC stream, or None when the stream is too short for
    pair-range shards to pay (hvb200_mc_segment_layout)."""
    p0, n = ctypes.c_uint64(), ctypes.c_uint64()
    rc = _load().hvb200_mc_segment_layout(nsamples, nobj, world, rank, ctypes.byref(p0), ctypes.byref(n))
    if rc < 0:
        raise ValueError(_last_error())
    return None if rc == 1 else (p0.value, n.value)


def mc_segments_resolve(segments, nobj: int, nsamples: int):
    """Per rank (pair, discard, i0, i1) from the gathered segment descriptions (hvb200_mc_segments_resolve)."""
    world = len(segments)
    segs = (McSegment * world)(*segments)
    starts = (McStart * world)()
    if _load().hvb200_mc_segments_resolve(segs, world, nobj, nsamples, starts) != 0:
        raise RuntimeError(f"hvb200_mc_segments_resolve failed: {_last_error()}")
    return [(st.pair, st.discard, st.i0, st.i1) for st in starts]


def hv_approx_range(points, ref, *, maximise=False, nsamples: int, seed: int = 0, method: str, i0: int, i1: int,
                    device: int = 0, stream: int | None = None):
    """One shard of an estimate: the (hi, lo) partial sum over samples [i0, i1) on `device`."""
    with Plan(points, ref, maximise, device=device) as plan:
        return plan.run(method, nsamples, seed, i0, i1, stream)


def hv_approx_shard(points, ref, *, maximise=False, nsamples: int, seed: int = 0, method: str, rank: int, world: int,
                    device: int = 0, stream: int | None = None, allow_extension: bool = False):
    """Shard `rank` of `world` of an estimate as a (hi, lo) partial sum: block-cyclic for DZ2019-HW (balanced),
    a contiguous range for the sequential MC and Rphi streams."""
    from .sharding import shard_block, shard_range
    with Plan(points, ref, maximise, device=device, allow_extension=allow_extension) as plan:
        if method == "DZ2019-HW" and world > 1:
            return plan.run_blocks(method, nsamples, rank, world, block=shard_block(nsamples, world), seed=seed, stream=stream)
        if method == "DZ2019-MC" and world > 1:
            from .sharding import mc_shard_start
            i0, i1 = mc_shard_start(plan, nsamples, seed, rank, world, device=device)
            return plan.run(method, nsamples, seed, i0, i1, stream)
        i0, i1 = shard_range(nsamples, rank, world)
        return plan.run(method, nsamples, seed, i0, i1, stream)


def hv_approx_finalize(nobj: int, nsamples: int, partials) -> float:
    """c_d * (sum of partials) / nsamples in long double (c/hvapprox.c:691-692)."""
    flat = np.ascontiguousarray(np.asarray(partials, dtype=np.float64).reshape(-1))
    return float(_load().hvb200_finalize(nobj, nsamples, _dp(flat), flat.size // 2))


def measure_fp64(device: int = 0) -> dict:
    dmul, dsetp = ctypes.c_double(), ctypes.c_double()
    flops = _load().hvb200_measure_fp64_flops(device, ctypes.byref(dmul), ctypes.byref(dsetp))
    return {"dfma_flops": flops, "dmul_per_s": dmul.value, "dsetp_per_s": dsetp.value}
