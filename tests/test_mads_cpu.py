"""CPU-only: the batched MADS poll driver's host logic on plain Python objectives."""
import numpy as np


def test_mads_minimises_quadratic_on_integer_mesh(cov):
    from coverage_b200 import mads
    target = np.array([37.0, -12.0, 250.0, 3.0])
    calls = {"batches": 0}

    class Obj:
        def __call__(self, x):
            return float(np.sum((x - target) ** 2))

        def batch(self, X):
            calls["batches"] += 1
            return np.sum((X - target) ** 2, axis=1)

    x0 = np.array([30.3, -5.2, 240.9, 10.1])
    res, runtime, st = mads.optimize(x0, Obj(), [], [], 200, seed=1, return_stats=True)
    assert runtime > 0 and st["batches"] == calls["batches"]
    assert np.all(res == np.rint(res))          # granularity 1.0 on every variable
    assert Obj()(res) <= 1.0 and Obj()(res) < Obj()(x0)
    assert st["evaluations"] <= 1 + 2 * 4 * st["iterations"]  # at most 2n trial points per poll


def test_mads_extreme_barrier_and_start_fallback(cov):
    from coverage_b200 import mads
    obj = lambda x: float(x[0] + x[1])  # noqa: E731  unbounded below: only the barrier stops it
    inside = lambda x: bool(np.hypot(x[0] - 10, x[1] - 10) <= 6)  # noqa: E731
    res, _ = mads.optimize(np.array([10.0, 10.0]), obj, [[lambda x: True, inside]], [], 100, seed=3)
    assert inside(res) and obj(res) < 20.0 - 6.0
    # nothing feasible anywhere: the start comes back (the reference returns p.i)
    res2, _ = mads.optimize(np.array([1.5, 2.5]), obj, [lambda x: False], [], 10, seed=3)
    assert res2.tolist() == [1.5, 2.5]


def test_mesh_ladder(cov):
    from coverage_b200.mads import _Mesh
    m = _Mesh(np.array([250.0, 12.0, 0.0]), 1.0)
    assert m.poll_size().tolist() == [20.0, 1.0, 1.0]
    m.enlarge()
    assert m.poll_size().tolist() == [50.0, 2.0, 2.0]
    for _ in range(10):
        last = m.refine()
    assert m.poll_size().tolist() == [1.0, 1.0, 1.0] and last is False
    assert m.mesh_size().tolist() == [1.0, 1.0, 1.0]


def test_multistart_lockstep_equals_single_solves(cov):
    """optimize_multistart batches the poll sets of S solves into one objective call per iteration; every solve's
    result is exactly what optimize() gives from the same start with the same random stream."""
    from coverage_b200 import mads
    centres = np.array([[40.0, 40.0, 10.0], [-30.0, 25.0, 4.0], [5.0, -60.0, 7.0]])
    calls = {"batches": 0, "rows": []}

    class Obj:  # three basins of different depth
        def batch(self, X):
            calls["batches"] += 1
            calls["rows"].append(len(X))
            d = ((X[:, None, :2] - centres[None, :, :2]) ** 2).sum(axis=2)
            return (d / 50.0 - centres[None, :, 2]).min(axis=1)

        def __call__(self, x):
            return float(self.batch(np.asarray(x, dtype=np.float64)[None, :])[0])

    ok = lambda x: bool(abs(x[0]) <= 80 and abs(x[1]) <= 80)  # noqa: E731
    starts = np.array([[30.5, 50.2], [-20.0, 20.0], [0.3, -40.9], [70.0, -70.0], [200.0, 200.0]])
    res, objs, runtime, st = mads.optimize_multistart(starts, Obj(), [ok], [], 150, seed=11, return_stats=True)
    assert res.shape == starts.shape and objs.shape == (5,) and runtime > 0
    assert st["batches"] == calls["batches"] <= st["iterations"] + 1      # one objective call per lockstep iteration
    assert max(calls["rows"]) > 2 * 2                                     # ... carrying several solves' poll sets
    assert objs[4] == np.inf and res[4].tolist() == [200.0, 200.0]       # a start with nothing feasible around it
    assert st["best"] == int(np.argmin(objs)) and objs.min() <= -9.9      # the deepest basin is found by some start
    for k in range(5):
        calls_before = calls["batches"]
        single, _, s1 = mads.optimize(starts[k], Obj(), [ok], [], 150, seed=11 + k, return_stats=True)
        assert single.tolist() == res[k].tolist(), k
        assert (s1["objective"] == objs[k]) or (not np.isfinite(objs[k]) and not np.isfinite(s1["objective"]))
        assert s1["iterations"] == st["per_solve_iterations"][k]
        assert calls["batches"] - calls_before == s1["batches"]
    import pytest
    with pytest.raises(ValueError):
        mads.optimize_multistart(np.zeros(3), Obj(), [], [], 5)
