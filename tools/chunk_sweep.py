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
for mb in (0, 4, 6, 8, 12, 16, 24, 32):
    e.set_option(cov.OPT_CHUNK, 0 if mb == 0 else int(mb * 2**20 // 120) // 32 * 32)
    for _ in range(3): e.eval_batch(X, out=out)
    t = time.perf_counter()
    for _ in range(30): e.eval_batch(X, out=out)
    print("slice MiB", mb if mb else "auto", "%.3f ms" % ((time.perf_counter() - t) / 30 * 1e3))
