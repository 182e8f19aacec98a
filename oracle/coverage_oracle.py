"""NumPy restatement of the reference's coverage-objective path, written independently of
coverage_oracle.c so the two can check each other.

TEST INFRASTRUCTURE ONLY -- nothing in the product imports this module (see oracle/README.md).

PARITY UNPINNED: the reference is Julia, Julia is not installed, and the reference's own test
suite (test/runtests.jl:1-6) holds no vectors for this path.  Every function cites the Julia
lines it restates; citations are relative to /root/reference/.

NumPy float64 ufuncs are IEEE-754 binary64, round-to-nearest, and `a*a + b*b` is evaluated as
separate multiply/add ufunc calls, so no FMA contraction can occur -- the same arithmetic as
Julia's `sqrt((px-cx)^2 + (py-cy)^2) < R`.
"""
from __future__ import annotations

import math

import numpy as np

TAN_HALF_FOV_DEFAULT = math.tan((100 / 180 * math.pi) / 2)  # src/FullSimulation.jl:735


def createPOI(dx: float, dy: float, x_length: float, y_length: float) -> np.ndarray:
    """src/AreaCoverageCalculation.jl:11-21 -- i outer, j inner; returns P x 5 float64."""
    i = np.arange(1.0, math.floor(x_length) + 1.0)
    j = np.arange(1.0, math.floor(y_length) + 1.0)
    px = i * dx - dx / 2
    py = j * dy - dy / 2
    pts = np.empty((len(i) * len(j), 5), dtype=np.float64)
    pts[:, 0] = np.repeat(px, len(j))
    pts[:, 1] = np.tile(py, len(i))
    pts[:, 2] = dx * dy
    pts[:, 3] = dx * dy
    pts[:, 4] = 0.0
    return pts


def make_circles(arr):
    """src/AreaCoverageCalculation.jl:33-45 -- [x;y;R] -> list of (x, y, R)."""
    arr = np.asarray(arr, dtype=np.float64)
    n = len(arr) // 3
    return [(arr[i], arr[n + i], arr[2 * n + i]) for i in range(n)]


def make_MADS(circles) -> np.ndarray:
    """src/AreaCoverageCalculation.jl:48-59 -- list of (x, y, R) -> [x;y;R]."""
    return np.array([c[0] for c in circles] + [c[1] for c in circles] + [c[2] for c in circles],
                    dtype=np.float64)


def covered_mask(circles, pts: np.ndarray) -> np.ndarray:
    """Boolean per list entry: covered by ANY disc (the first-hit `break` only decides which
    disc claims the point, not whether it is counted).  Predicate of
    src/AreaCoverageCalculation.jl:70."""
    circles = np.asarray(circles, dtype=np.float64)
    n = len(circles) // 3
    px = pts[:, 0]
    py = pts[:, 1]
    cov = np.zeros(len(pts), dtype=bool)
    for c in range(n):
        ddx = px - circles[c]
        ddy = py - circles[n + c]
        s = ddx * ddx + ddy * ddy
        with np.errstate(invalid="ignore"):
            cov |= np.sqrt(s) < circles[2 * n + c]
    return cov


def calculateArea(circles, pts: np.ndarray):
    """src/AreaCoverageCalculation.jl:63-110 -- returns (area, count).  The area is the
    sequential Float64 sum of the covered entries' weights in list order (np.cumsum is a
    strictly sequential accumulation, unlike np.sum's pairwise scheme)."""
    cov = covered_mask(circles, pts)
    w = pts[cov, 3]
    area = float(np.cumsum(w)[-1]) if len(w) else 0.0
    return area, int(cov.sum())


def objective(x, pts: np.ndarray, N: int, r_max):
    """src/TDM_STATIC_opt.jl:82-100 -- returns (objective, count)."""
    x = np.asarray(x, dtype=np.float64)
    r_max = np.asarray(r_max, dtype=np.float64)
    area, count = calculateArea(x, pts)
    violation = 0.0
    for i in range(N):
        violation += abs(float(x[i + 2 * N]) - float(r_max[i]))
    return -area + violation * 1e5, count


def cons3(x, pre, tan_half_fov: float, d_lim) -> bool:
    """src/TDM_Constraints.jl:54-75."""
    x = np.asarray(x, dtype=np.float64)
    pre = np.asarray(pre, dtype=np.float64)
    n = len(x) // 3
    for i in range(n):
        z1 = float(pre[2 * n + i]) / tan_half_fov
        z2 = float(x[2 * n + i]) / tan_half_fov
        ax = float(pre[i]) - float(x[i])
        ay = float(pre[n + i]) - float(x[n + i])
        az = z1 - z2
        if math.sqrt(ax * ax + ay * ay + az * az) > float(d_lim[i]):
            return False
    return True


def cons7(x, tan_half_fov: float) -> bool:
    """src/TDM_Constraints.jl:142-154."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x) // 3
    for i in range(n):
        if x[i + n] < 200 and x[i + 2 * n] > 19 * tan_half_fov:
            return False
    return True


def cons8(x, sep: float = 15.0) -> bool:
    """src/TDM_Constraints.jl:157-172."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x) // 3
    for i in range(n):
        for j in range(n):
            if j != i:
                ax = float(x[i]) - float(x[j])
                ay = float(x[i + n]) - float(x[j + n])
                if math.sqrt(ax * ax + ay * ay) < sep:
                    return False
    return True


def _julia_max0(v: float) -> float:
    """Julia's max(v, 0.0) on Float64: NaN propagates, max(-0.0, 0.0) is +0.0."""
    if v != v:
        return v
    return v if v > 0.0 else 0.0


def cons1_progressive(x, r_max) -> float:
    """src/TDM_Constraints.jl:182-195."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x) // 3
    violation = 0.0
    for i in range(n):
        violation += _julia_max0(float(x[2 * n + i]) - float(r_max[i]))
    return violation


def consK_progressive(x, r_max, which: int) -> float:
    """src/TDM_Constraints.jl:197-221 -- cons2_progressive (which = 2) / cons3_progressive (which = 3):
    the same term for ONE fixed UAV, 1-based index as in the reference."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x) // 3
    violation = 0
    violation += _julia_max0(float(x[2 * n + which - 1]) - float(r_max[which - 1]))
    return violation


def rmvCoveredPOI(circles, pts: np.ndarray) -> np.ndarray:
    """src/CellFunctions.jl:81-108 -- returns the surviving points in list order."""
    return pts[~covered_mask(circles, pts)]


def allocate_even_circles(r_centering_cir: float, N: int, r_uav: float, center_x: float,
                          center_y: float) -> np.ndarray:
    """src/Base_Functions.jl:44-65."""
    xs, ys, rs = [], [], []
    for i in range(1, N + 1):
        ref_angle = 2 * math.pi / N * (i - 1)
        xs.append(r_centering_cir * math.cos(ref_angle) + center_x)
        ys.append(r_centering_cir * math.sin(ref_angle) + center_y)
        rs.append(r_uav)
    return np.array(xs + ys + rs, dtype=np.float64)


def threshold_by_search(R: float) -> float:
    """T(R) = smallest double t with sqrt(t) >= R (see coverage_oracle.c)."""
    if not (R > 0):
        return 0.0
    if math.isinf(R):
        return math.inf
    t = R * R
    if math.isinf(t):
        t = 1.7976931348623157e308
        if math.sqrt(t) < R:
            return math.inf
    while t > 0 and math.sqrt(t) >= R:
        t = math.nextafter(t, -math.inf)
    while math.sqrt(t) < R:
        t = math.nextafter(t, math.inf)
    return t


# ---- continuous variant: exact area of the union of discs -------------------------
# PARITY UNPINNED: the reference ships only the circle-circle primitives of a Green's-theorem method
# (src/Base_Functions.jl:230-355: distance, contained, intersection) and no driver, so there is nothing
# to compare against except mathematics: closed forms for <= 2 discs and Monte-Carlo estimates.
def union_area(x) -> float:
    """Area of the union of the discs [x;y;R] by boundary integration: for every circle the arcs not
    inside any other disc contribute 1/2 * integral (x dy - y dx)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x) // 3
    cx, cy, R = x[:n], x[n:2 * n], x[2 * n:]
    total = 0.0
    two_pi = 2.0 * math.pi
    for i in range(n):
        if not (R[i] > 0):
            continue
        covered = []  # angular intervals of circle i inside some other disc
        whole = False
        for j in range(n):
            if j == i or not (R[j] > 0):
                continue
            dx, dy = cx[j] - cx[i], cy[j] - cy[i]
            d = math.hypot(dx, dy)
            if d >= R[i] + R[j]:
                continue
            if d + R[i] <= R[j]:  # circle i inside disc j (identical discs: the lower index survives)
                if d + R[j] <= R[i] and i < j:
                    continue
                whole = True
                break
            if d + R[j] <= R[i]:  # disc j inside disc i: does not touch i's boundary
                continue
            phi = math.atan2(dy, dx)
            c = (R[i] * R[i] + d * d - R[j] * R[j]) / (2.0 * R[i] * d)
            alpha = math.acos(max(-1.0, min(1.0, c)))
            a = (phi - alpha) % two_pi
            b = a + 2.0 * alpha
            if b > two_pi:
                covered.append((a, two_pi))
                covered.append((0.0, b - two_pi))
            else:
                covered.append((a, b))
        if whole:
            continue
        covered.sort()
        pos = 0.0
        arcs = []
        for a, b in covered:
            if a > pos:
                arcs.append((pos, a))
            pos = max(pos, b)
        if pos < two_pi:
            arcs.append((pos, two_pi))
        for t1, t2 in arcs:
            total += 0.5 * (R[i] * R[i] * (t2 - t1) +
                            R[i] * (cx[i] * (math.sin(t2) - math.sin(t1)) - cy[i] * (math.cos(t2) - math.cos(t1))))
    return total
