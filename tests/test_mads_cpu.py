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
