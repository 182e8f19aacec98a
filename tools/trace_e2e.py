"""Print the device timeline of one cov_eval_batch call on the bench workload (debug aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
bits, n = cov.synth.fire_grid(256); d = 500 / 256
e.set_grid_bits(bits, 256, 256, d, d); e.set_params(5, np.full(5, 30 * cov.TAN_HALF_FOV_DEFAULT))
B = 1_000_000
X = e.pinned((B, 15)); cov.synth.random_candidates(B, 5, seed=3, out=X)
out = {"obj": e.pinned((B,)), "count": e.pinned((B,), np.int64), "feasible": e.pinned((B,), np.uint8)}
for _ in range(3): e.eval_batch(X, out=out)
e.set_option(cov.OPT_TRACE, 1)
t = time.perf_counter(); e.eval_batch(X, out=out); dt = time.perf_counter() - t
print("wall ms", dt * 1e3)
np.set_printoptions(precision=3, suppress=True, linewidth=150)
print(e.trace())
e.set_option(cov.OPT_TRACE, 0)
t = time.perf_counter()
for _ in range(20): e.eval_batch(X, out=out)
print("avg wall ms (no trace)", (time.perf_counter() - t) / 20 * 1e3)
