"""e2e time of cov_eval_batch with results written straight to pinned host memory vs copied back."""
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
Xu = np.array(X)
for zc in (1, 0, 1, 0):
    e.set_option(cov.OPT_ZEROCOPY_OUT, zc)
    for _ in range(3): e.eval_batch(X, out=out)
    t = time.perf_counter()
    for _ in range(30): e.eval_batch(X, out=out)
    dt = (time.perf_counter() - t) / 30
    r = e.eval_batch(Xu)
    t = time.perf_counter()
    for _ in range(10): e.eval_batch(Xu, out=r)
    du = (time.perf_counter() - t) / 10
    print("zero-copy out =", zc, ": pinned %.3f ms, pageable %.3f ms per 1M" % (dt * 1e3, du * 1e3))
