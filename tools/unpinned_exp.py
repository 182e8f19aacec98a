"""Time cov_eval_batch on pageable host buffers (what a Julia caller passes) vs pinned ones."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
bits, n = cov.synth.fire_grid(256); d = 500 / 256
e.set_grid_bits(bits, 256, 256, d, d); e.set_params(5, np.full(5, 30 * cov.TAN_HALF_FOV_DEFAULT))
B = 1_000_000
X = cov.synth.random_candidates(B, 5, seed=3)
for _ in range(3): r = e.eval_batch(X)
t = time.perf_counter()
for _ in range(10): r = e.eval_batch(X, out=r)
print("pageable in/out: %.2f ms per 1M" % ((time.perf_counter() - t) / 10 * 1e3))
Xp = e.pinned((B, 15)); Xp[:] = X
out = {"obj": e.pinned((B,)), "count": e.pinned((B,), np.int64), "feasible": e.pinned((B,), np.uint8)}
for _ in range(3): e.eval_batch(Xp, out=out)
t = time.perf_counter()
for _ in range(10): e.eval_batch(Xp, out=out)
print("pinned in/out:   %.2f ms per 1M" % ((time.perf_counter() - t) / 10 * 1e3))
x = X[0].copy()
for _ in range(10): e.eval_one(x)
t = time.perf_counter()
for _ in range(1000): e.eval_one(x)
print("eval_one: %.1f us per call" % ((time.perf_counter() - t) / 1000 * 1e6))
P = X[:30].copy()
for _ in range(10): e.eval_batch(P)
t = time.perf_counter()
for _ in range(1000): e.eval_batch(P)
print("eval_batch(30 candidates, a MADS poll): %.1f us per call" % ((time.perf_counter() - t) / 1000 * 1e6))
