"""ctypes wrapper of oracle/libcoverage_oracle.so (coverage_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "libcoverage_oracle.so")
SRC = os.path.join(_HERE, "coverage_oracle.c")


def build(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([gcc, "-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-fast-math",
                               "-fvisibility=hidden", "-o", SO, SRC, "-lm"])
    return SO


_vp, _i64, _d, _i = C.c_void_p, C.c_int64, C.c_double, C.c_int
_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.orc_createPOI.restype = _i64
        L.orc_createPOI.argtypes = [_d, _d, _d, _d, _vp]
        L.orc_calculateArea.restype = _d
        L.orc_calculateArea.argtypes = [_vp, _i64, _vp, _i64, C.POINTER(_i64), C.POINTER(_i64)]
        L.orc_objective.restype = _d
        L.orc_objective.argtypes = [_vp, _i64, _vp, _vp, _i64, C.POINTER(_i64)]
        L.orc_cons3.restype = _i
        L.orc_cons3.argtypes = [_vp, _i64, _vp, _d, _vp]
        L.orc_cons7.restype = _i
        L.orc_cons7.argtypes = [_vp, _i64, _d]
        L.orc_cons8.restype = _i
        L.orc_cons8.argtypes = [_vp, _i64, _d]
        L.orc_cons1_progressive.restype = _d
        L.orc_cons1_progressive.argtypes = [_vp, _i64, _vp]
        L.orc_consK_progressive.restype = _d
        L.orc_consK_progressive.argtypes = [_vp, _i64, _vp, _i64]
        L.orc_rmvCoveredPOI.restype = _i64
        L.orc_rmvCoveredPOI.argtypes = [_vp, _i64, _vp, _i64]
        L.orc_allocate_even_circles.restype = None
        L.orc_allocate_even_circles.argtypes = [_d, _i64, _d, _d, _d, _vp]
        L.orc_eval_batch.restype = None
        L.orc_eval_batch.argtypes = [_vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _d, _d, _i, _vp, _vp, _vp, _vp, _i]
        L.orc_class_counts.restype = None
        L.orc_class_counts.argtypes = [_vp, _i64, _vp, _i64, _vp, _i64, _vp]
        L.orc_fire_step_philox.restype = _i64
        L.orc_fire_step_philox.argtypes = [_vp, _vp, _i64, _i64, _d, _d, _d, _d, _d, _i64, C.c_uint64, _vp]
        L.orc_threshold_by_search.restype = _d
        L.orc_threshold_by_search.argtypes = [_d]
        L.orc_num_threads.restype = _i
        L.orc_num_threads.argtypes = []
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def createPOI(dx, dy, x_length, y_length) -> np.ndarray:
    n = lib().orc_createPOI(dx, dy, x_length, y_length, None)
    out = np.empty((n, 5), dtype=np.float64)
    lib().orc_createPOI(dx, dy, x_length, y_length, _p(out))
    return out


def calculateArea(circles, pts5):
    circles, pts5 = _f(circles).ravel(), _f(pts5).reshape(-1, 5)
    cnt, tests = C.c_int64(), C.c_int64()
    a = lib().orc_calculateArea(_p(circles), circles.size // 3, _p(pts5), pts5.shape[0], C.byref(cnt), C.byref(tests))
    return a, cnt.value, tests.value


def objective(x, r_max, pts5):
    x, r_max, pts5 = _f(x).ravel(), _f(r_max).ravel(), _f(pts5).reshape(-1, 5)
    cnt = C.c_int64()
    v = lib().orc_objective(_p(x), x.size // 3, _p(r_max), _p(pts5), pts5.shape[0], C.byref(cnt))
    return v, cnt.value


def cons3(x, pre, tan_half_fov, d_lim) -> bool:
    x, pre, d_lim = _f(x).ravel(), _f(pre).ravel(), _f(d_lim).ravel()
    return bool(lib().orc_cons3(_p(x), x.size // 3, _p(pre), tan_half_fov, _p(d_lim)))


def cons7(x, tan_half_fov) -> bool:
    x = _f(x).ravel()
    return bool(lib().orc_cons7(_p(x), x.size // 3, tan_half_fov))


def cons8(x, sep=15.0) -> bool:
    x = _f(x).ravel()
    return bool(lib().orc_cons8(_p(x), x.size // 3, sep))


def cons1_progressive(x, r_max) -> float:
    x, r_max = _f(x).ravel(), _f(r_max).ravel()
    return lib().orc_cons1_progressive(_p(x), x.size // 3, _p(r_max))


def consK_progressive(x, r_max, which: int) -> float:
    """cons2_progressive (which = 2) / cons3_progressive (which = 3), src/TDM_Constraints.jl:197-221."""
    x, r_max = _f(x).ravel(), _f(r_max).ravel()
    return lib().orc_consK_progressive(_p(x), x.size // 3, _p(r_max), which)


def rmvCoveredPOI(circles, pts5) -> np.ndarray:
    circles = _f(circles).ravel()
    pts = _f(pts5).reshape(-1, 5).copy()
    keep = lib().orc_rmvCoveredPOI(_p(circles), circles.size // 3, _p(pts), pts.shape[0])
    return pts[:keep]


def allocate_even_circles(r_centering_cir, N, r_uav, center_x, center_y) -> np.ndarray:
    out = np.empty(3 * N, dtype=np.float64)
    lib().orc_allocate_even_circles(r_centering_cir, N, r_uav, center_x, center_y, _p(out))
    return out


def eval_batch(X, N, r_max, pts5, pre=None, d_lim=None, tan_half_fov=0.0, sep_min=0.0, use_cons7=False,
               want_prog=False, threads=0):
    X, r_max, pts5 = _f(X).reshape(-1, 3 * N), _f(r_max).ravel(), _f(pts5).reshape(-1, 5)
    B = X.shape[0]
    obj = np.empty(B, dtype=np.float64)
    count = np.empty(B, dtype=np.int64)
    feas = np.empty(B, dtype=np.uint8)
    prog = np.empty(B, dtype=np.float64) if want_prog else None
    pre = None if pre is None else _f(pre).ravel()
    d_lim = None if d_lim is None else _f(np.broadcast_to(np.asarray(d_lim, dtype=np.float64), (N,)))
    lib().orc_eval_batch(_p(X), B, N, _p(r_max), _p(pts5), pts5.shape[0], _p(pre), _p(d_lim), tan_half_fov,
                         sep_min, 1 if use_cons7 else 0, _p(obj), _p(count), _p(feas), _p(prog), threads)
    return {"obj": obj, "count": count, "feasible": feas, "progressive": prog}


def class_counts(circles, pts5, class_of, n_classes):
    circles, pts5 = _f(circles).ravel(), _f(pts5).reshape(-1, 5)
    class_of = np.ascontiguousarray(class_of, dtype=np.int32)
    out = np.zeros(n_classes, dtype=np.int64)
    lib().orc_class_counts(_p(circles), circles.size // 3, _p(pts5), pts5.shape[0], _p(class_of), n_classes, _p(out))
    return out


def fire_step(grid, nx, ny, dx, dy, wind_speed, wind_direction, prob_spread, step, seed):
    """One update_grid() step; grid: nx*ny uint8 (index (i-1) + nx*(j-1)). Returns (new_grid, pts5)."""
    grid = np.ascontiguousarray(grid, dtype=np.uint8).ravel()
    new = np.empty_like(grid)
    n = lib().orc_fire_step_philox(_p(grid), _p(new), nx, ny, dx, dy, wind_speed, wind_direction, prob_spread,
                                   step, seed, None)
    pts = np.empty((n, 5), dtype=np.float64)
    lib().orc_fire_step_philox(_p(grid), _p(new), nx, ny, dx, dy, wind_speed, wind_direction, prob_spread,
                               step, seed, _p(pts))
    return new, pts


def threshold_by_search(R: float) -> float:
    return lib().orc_threshold_by_search(float(R))


def num_threads() -> int:
    return lib().orc_num_threads()
