"""Host-side mirror of the reference module `AreaCoverageCalculation`
(/root/reference/src/AreaCoverageCalculation.jl): same function names, argument meaning and
results; the cell-vs-disc arithmetic runs in libcoverage_cuda's kernels.

A point list is the reference's `Vector{Vector{Float64}}` of `[x, y, area, weight, covered]`
entries, held here as a `PointList` (a P x 5 float64 array plus the lattice it lives on).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .engine import CoverageEngine


@dataclass
class Circle:
    """src/Base_Functions.jl:37-41."""
    x: float
    y: float
    R: float


class PointList:
    """P x 5 float64 `[x, y, area, weight, covered]` on an nx x ny lattice of dx x dy cells
    (cell (i, j), 1-based, has its centre at (i*dx - dx/2, j*dy - dy/2),
    src/AreaCoverageCalculation.jl:16).  `version` changes whenever the list does, so device
    copies know when to refresh."""

    def __init__(self, data, nx: int, ny: int, dx: float, dy: float):
        self.data = np.ascontiguousarray(data, dtype=np.float64).reshape(-1, 5)
        self.nx, self.ny, self.dx, self.dy = int(nx), int(ny), float(dx), float(dy)
        self.version = 0

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, k):
        return self.data[k]

    def __iter__(self):
        return iter(self.data)

    def cell_index(self) -> np.ndarray:
        """0-based (i-1) + nx*(j-1) of every entry."""
        i = np.rint((self.data[:, 0] + self.dx / 2) / self.dx).astype(np.int64)
        j = np.rint((self.data[:, 1] + self.dy / 2) / self.dy).astype(np.int64)
        return (i - 1) + self.nx * (j - 1)

    def append(self, pts5):
        pts5 = np.asarray(pts5, dtype=np.float64).reshape(-1, 5)
        self.data = np.concatenate([self.data, pts5], axis=0)
        self.version += 1

    def keep(self, mask: np.ndarray):
        self.data = np.ascontiguousarray(self.data[mask])
        self.version += 1

    @staticmethod
    def infer(points) -> "PointList":
        """Accept a bare P x 5 array / list of 5-vectors: dx, dy from the `area` column and the
        coordinate spacing, nx, ny from the extent."""
        if isinstance(points, PointList):
            return points
        a = np.asarray(points, dtype=np.float64).reshape(-1, 5)
        if a.shape[0] == 0:
            return PointList(a, 1, 1, 1.0, 1.0)

        def spacing(v):
            u = np.unique(v)
            if u.size > 1:
                return float(np.min(np.diff(u)))
            return None
        dx, dy = spacing(a[:, 0]), spacing(a[:, 1])
        area = float(a[0, 2])
        if dx is None and dy is None:
            dx = dy = math.sqrt(area) if area > 0 else 1.0
        elif dx is None:
            dx = area / dy if area > 0 else dy
        elif dy is None:
            dy = area / dx if area > 0 else dx
        nx = int(round((float(a[:, 0].max()) + dx / 2) / dx))
        ny = int(round((float(a[:, 1].max()) + dy / 2) / dy))
        return PointList(a, max(nx, 1), max(ny, 1), dx, dy)


def createPOI(dx: float, dy: float, x_length: float, y_length: float) -> PointList:
    """src/AreaCoverageCalculation.jl:11-21 -- i outer, j inner, centre (i*dx - dx/2, j*dy - dy/2),
    area = weight = dx*dy, covered = false."""
    nx, ny = int(math.floor(x_length)), int(math.floor(y_length))
    i = np.arange(1.0, nx + 1.0)
    j = np.arange(1.0, ny + 1.0)
    pts = np.empty((nx * ny, 5), dtype=np.float64)
    pts[:, 0] = np.repeat(i * dx - dx / 2, ny)
    pts[:, 1] = np.tile(j * dy - dy / 2, nx)
    pts[:, 2] = dx * dy
    pts[:, 3] = dx * dy
    pts[:, 4] = 0.0
    return PointList(pts, nx, ny, dx, dy)


def make_circles(arr):
    """src/AreaCoverageCalculation.jl:33-45 -- [x;y;R] -> list of Circle."""
    arr = np.asarray(arr, dtype=np.float64).ravel()
    n = arr.size // 3
    return [Circle(float(arr[i]), float(arr[n + i]), float(arr[2 * n + i])) for i in range(n)]


def make_MADS(circles) -> np.ndarray:
    """src/AreaCoverageCalculation.jl:48-59 -- list of Circle -> [x;y;R]."""
    return np.array([c.x for c in circles] + [c.y for c in circles] + [c.R for c in circles], dtype=np.float64)


class ResidentList:
    """A PointList with its device-resident copy (one CoverageEngine)."""

    def __init__(self, points: PointList, device: int = 0, engine: CoverageEngine | None = None):
        self.points = points
        self.engine = engine if engine is not None else CoverageEngine(device)
        self._synced = None
        self._cell_index = None

    def sync(self):
        p = self.points
        if self._synced != (id(p), p.version):
            self.engine.set_points(p.data, p.nx, p.ny, p.dx, p.dy)
            self._synced = (id(p), p.version)
            self._cell_index = None
        return self.engine

    def mark_synced(self):
        """The caller changed host list and device store consistently by itself."""
        self._synced = (id(self.points), self.points.version)
        self._cell_index = None

    def cell_index(self):
        if self._cell_index is None:
            self._cell_index = self.points.cell_index()
        return self._cell_index

    def area_and_count(self, circles) -> tuple[float, int]:
        """calculateArea's result and the covered-entry count for ONE disc vector."""
        eng = self.sync()
        circles = np.ascontiguousarray(circles, dtype=np.float64).ravel()
        n = circles.size // 3
        # the objective with the penalty switched off is -area exactly: -area + violation * 0.0 (finite radii; a
        # non-finite radius makes the product NaN, as it makes the reference's own objective)
        if eng.N != n or eng.param_owner != ("calculateArea", n):
            eng.set_params(n, np.zeros(n), penalty_scale=0.0)
            eng.param_owner = ("calculateArea", n)
        res = eng.eval_batch(circles.reshape(1, -1), want_feasible=False)
        # dyadic weights: sum_k w_k * count_k in any order; otherwise the library replays the reference's list-order
        # sum on the device (ordered kernel) -- either way this is calculateArea's Float64
        area = 0.0 - float(res["obj"][0])
        return (area if area == area else self._area_on_host(circles)), int(res["count"][0])

    def _area_on_host(self, circles) -> float:
        """Non-finite radii only: replay the ordered sum from the device's covered mask."""
        covered = self.engine.covered_mask(circles)[self.cell_index()].astype(bool)
        w = self.points.data[covered, 3]
        return float(np.cumsum(w)[-1]) if w.size else 0.0  # cumsum: strictly sequential


_resident_cache: dict[int, ResidentList] = {}


def _resident(points) -> ResidentList:
    if isinstance(points, ResidentList):
        return points
    pl = PointList.infer(points)
    key = id(points)
    r = _resident_cache.get(key)
    if r is None or r.points is not pl and not isinstance(points, PointList):
        if len(_resident_cache) > 8:
            _resident_cache.clear()
        r = ResidentList(pl)
        if isinstance(points, PointList):
            _resident_cache[key] = r
    return r


def calculateArea(circles, points) -> float:
    """src/AreaCoverageCalculation.jl:63-110 -- sum of the weights (entry 4) of the list entries
    covered by at least one disc, `sqrt((px-cx)^2 + (py-cy)^2) < R` in Float64, strict <."""
    return _resident(points).area_and_count(circles)[0]


def rmvCoveredPOI(circles, points):
    """src/AreaCoverageCalculation.jl:113-137 -- delete the covered entries, keep list order."""
    r = _resident(points)
    eng = r.sync()
    circles = np.ascontiguousarray(circles, dtype=np.float64).ravel()
    n = circles.size // 3
    if eng.N != n:
        eng.set_params(n, np.zeros(n), penalty_scale=0.0)
    covered = eng.covered_mask(circles)[r.cell_index()].astype(bool)
    eng.remove_covered(circles)
    r.points.keep(~covered)
    r.mark_synced()
    return r.points


_union_engine = None


def unionArea(circles, engine: CoverageEngine | None = None) -> float:
    """Continuous variant: exact area of the union of the discs [x;y;R], by boundary
    integration on the device (cov_union_area_batch).  The reference ships only the pair primitives of
    this method (src/Base_Functions.jl:230-355) and no driver."""
    global _union_engine
    circles = np.ascontiguousarray(circles, dtype=np.float64).ravel()
    if engine is None:
        if _union_engine is None:
            _union_engine = CoverageEngine(0)
        engine = _union_engine
    return float(engine.union_area(circles.reshape(1, -1), circles.size // 3)[0])
