"""CPU-only: the oracle (C and NumPy restatements) against the golden vectors and each other."""
import math

import numpy as np

T = math.tan((100 / 180 * math.pi) / 2)


def test_kat1_createPOI(orc, npo, kat):
    pts = orc.createPOI(5.0, 5.0, 100.0, 100.0)
    assert len(pts) == kat["kat1"]["P"]
    assert pts[0].tolist() == kat["kat1"]["first"]
    assert pts[1].tolist() == kat["kat1"]["second"]  # j is the inner loop
    assert pts[-1].tolist() == kat["kat1"]["last"]
    assert np.array_equal(pts, npo.createPOI(5.0, 5.0, 100.0, 100.0))


def test_kat2_single_discs(orc, npo, kat):
    pts = orc.createPOI(5.0, 5.0, 100.0, 100.0)
    for k in kat["kat2"]:
        area, count, tests = orc.calculateArea(k["disc"], pts)
        assert (count, area) == (k["count"], k["area"])
        assert npo.calculateArea(np.array(k["disc"]), pts) == (k["area"], k["count"])


def test_tie_is_not_covered(orc):
    # cell (2.5, 2.5) is at distance exactly 5 from (5.5, 6.5): strict < must reject it
    pts = np.array([[2.5, 2.5, 25.0, 25.0, 0.0]])
    assert orc.calculateArea([5.5, 6.5, 5.0], pts)[1] == 0
    assert orc.calculateArea([5.5, 6.5, math.nextafter(5.0, 6.0)], pts)[1] == 1


def test_kat3_kat4_objective(orc, npo, kat):
    pts = orc.createPOI(5.0, 5.0, 100.0, 100.0)
    x3 = orc.allocate_even_circles(15.0, 5, 10 * T, 250.0, 250.0)
    assert x3.tolist() == kat["kat3"]["x"]
    for name in ("kat3", "kat4"):
        k = kat[name]
        obj, count = orc.objective(k["x"], k["r_max"], pts)
        assert (obj, count) == (k["objective"], k["count"])
        assert npo.objective(np.array(k["x"]), pts, 5, np.array(k["r_max"])) == (k["objective"], k["count"])
    assert orc.calculateArea(kat["kat3"]["x"], pts)[2] == 49846  # predicate evaluations incl. early break


def test_kat5_fire_list(orc, fire_rows, kat):
    k = kat["kat5"]
    allp = np.concatenate(fire_rows)
    assert (len(fire_rows), len(allp)) == (k["rows"], k["entries"])
    assert len(np.concatenate(fire_rows[:10])) == k["entries_10"]
    area, count, _ = orc.calculateArea(k["all_covering_disc"], allp)
    assert count == k["entries"] and area == 25.0 * k["entries"]  # list entries, not unique cells
    assert len(np.unique(allp[:, :2], axis=0)) == k["unique"]


def test_c_vs_numpy_random(orc, npo):
    rng = np.random.default_rng(7)
    pts = orc.createPOI(5.0, 5.0, 100.0, 100.0)
    r_max = np.full(4, 30 * T)
    for _ in range(25):
        x = np.concatenate([rng.random(8) * 500, (5 + rng.random(4) * 25) * T])
        assert orc.objective(x, r_max, pts) == npo.objective(x, pts, 4, r_max)
        pre = x + rng.normal(0, 4, 12)
        assert orc.cons3(x, pre, T, np.full(4, 10.0)) == npo.cons3(x, pre, T, np.full(4, 10.0))
        assert orc.cons7(x, T) == npo.cons7(x, T)
        assert orc.cons8(x) == npo.cons8(x)
        assert orc.cons1_progressive(x, r_max) == npo.cons1_progressive(x, r_max)
        assert np.array_equal(orc.rmvCoveredPOI(x, pts), npo.rmvCoveredPOI(x, pts))


def test_progressive_julia_max_semantics(orc, npo):
    """src/TDM_Constraints.jl:182-221: Julia's max(v, 0.0) propagates NaN (C's fmax would not) and
    max(-0.0, 0.0) is +0.0; both restatements must agree bit for bit, for cons1/2/3_progressive."""
    nan, inf = math.nan, math.inf
    r_max = np.array([10.0, 20.0, 30.0, inf])
    rows = [
        [0, 0, 0, 0, 0, 0, 0, 0, 12.0, 19.0, nan, 1.0],    # NaN radius: every sum it enters is NaN
        [0, 0, 0, 0, 0, 0, 0, 0, nan, 25.0, 31.0, 1.0],
        [0, 0, 0, 0, 0, 0, 0, 0, inf, 20.0, 30.0, inf],    # inf - inf = NaN for UAV 4
        [0, 0, 0, 0, 0, 0, 0, 0, 10.0, 20.0, 30.0, -inf],  # all differences <= 0 (one is -0.0 + ...)
        [0, 0, 0, 0, 0, 0, 0, 0, 10.0, 20.5, 29.0, 5.0],
        [0, 0, 0, 0, 0, 0, 0, 0, -0.0, 1e308, 1e-320, 0.0],
    ]
    for x in rows:
        x = np.array(x)
        a, b = orc.cons1_progressive(x, r_max), npo.cons1_progressive(x, r_max)
        assert np.float64(a).view(np.uint64) == np.float64(b).view(np.uint64) or (a != a and b != b), (x, a, b)
        for which in (1, 2, 3, 4):
            a, b = orc.consK_progressive(x, r_max, which), npo.consK_progressive(x, r_max, which)
            assert np.float64(a).view(np.uint64) == np.float64(b).view(np.uint64) or (a != a and b != b), (x, which)
    assert math.isnan(orc.cons1_progressive(np.array(rows[0]), r_max))
    assert math.isnan(orc.consK_progressive(np.array(rows[0]), r_max, 3))
    assert orc.consK_progressive(np.array(rows[0]), r_max, 2) == 0.0
    assert orc.consK_progressive(np.array(rows[4]), r_max, 2) == 0.5
    assert math.copysign(1.0, orc.consK_progressive(np.array(rows[3]), r_max, 1)) == 1.0  # +0.0, not -0.0
    out = orc.eval_batch(np.array(rows), 4, r_max, orc.createPOI(5.0, 5.0, 4.0, 4.0), want_prog=True)
    assert np.isnan(out["progressive"][:3]).all() and out["progressive"][4] == 0.5


def test_batch_matches_scalar_and_threads(orc):
    rng = np.random.default_rng(11)
    pts = orc.createPOI(5.0, 5.0, 40.0, 40.0)
    N, B = 3, 200
    X = np.concatenate([rng.random((B, 2 * N)) * 200, (5 + rng.random((B, N)) * 25) * T], axis=1)
    r_max = np.full(N, 30 * T)
    pre = X[0] + 1.0
    one = orc.eval_batch(X, N, r_max, pts, pre=pre, d_lim=10.0, tan_half_fov=T, sep_min=15.0, use_cons7=True,
                         want_prog=True, threads=1)
    many = orc.eval_batch(X, N, r_max, pts, pre=pre, d_lim=10.0, tan_half_fov=T, sep_min=15.0, use_cons7=True,
                          want_prog=True, threads=4)
    for k in one:
        assert np.array_equal(one[k], many[k])
    for b in range(0, B, 17):
        obj, cnt = orc.objective(X[b], r_max, pts)
        assert (one["obj"][b], one["count"][b]) == (obj, cnt)
        ok = orc.cons3(X[b], pre, T, np.full(N, 10.0)) and orc.cons8(X[b]) and orc.cons7(X[b], T)
        assert bool(one["feasible"][b]) == ok


def test_threshold_definition(orc, npo):
    rng = np.random.default_rng(3)
    for R in [5.0, 36.0, 11.9175359259421, 1.0, 2.0, 0.5, 1e-3, 1e6, 15.0] + list(rng.random(50) * 40):
        t = orc.threshold_by_search(R)
        assert t == npo.threshold_by_search(R)
        assert math.sqrt(t) >= R and math.sqrt(math.nextafter(t, 0.0)) < R
    assert orc.threshold_by_search(0.0) == 0.0 and orc.threshold_by_search(-1.0) == 0.0
    assert orc.threshold_by_search(float("nan")) == 0.0 and orc.threshold_by_search(math.inf) == math.inf


def test_union_area_closed_forms(npo):
    """Continuous variant (unpinned by the reference: checked against mathematics)."""
    lens = 2 * math.pi - (2 * math.acos(0.5) - 0.5 * math.sqrt(3.0))
    assert abs(npo.union_area([0, 0, 1.0]) - math.pi) < 1e-14
    assert abs(npo.union_area([0, 1, 0, 0, 1, 1.0]) - lens) < 1e-13
    assert abs(npo.union_area([0, 0.1, 0, 0, 2, 1.0]) - 4 * math.pi) < 1e-13
    assert abs(npo.union_area([0, 0, 0, 0, 1, 1.0]) - math.pi) < 1e-14
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(16) * 50, rng.random(8) * 10 + 2])
    pts = rng.random((400000, 2)) * 70 - 10
    cov = np.zeros(len(pts), dtype=bool)
    for i in range(8):
        cov |= (pts[:, 0] - x[i]) ** 2 + (pts[:, 1] - x[8 + i]) ** 2 < x[16 + i] ** 2
    assert abs(npo.union_area(x) - cov.mean() * 4900) < 0.01 * npo.union_area(x)  # Monte-Carlo, 1 %
