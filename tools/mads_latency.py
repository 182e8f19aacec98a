import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
T = cov.TAN_HALF_FOV_DEFAULT
e.set_grid_full(100, 100, 5.0, 5.0)
N = 5
e.set_params(N, np.full(N, 30 * T))
rng = np.random.default_rng(0)
x0 = np.concatenate([rng.random(N) * 500, rng.random(N) * 500, np.full(N, 20 * T)])
for zc in (1, 0):
    e.set_option(cov.OPT_ZEROCOPY_OUT, zc)
    for _ in range(3): r = e.mads_solve(x0, 100, 1.0, seed=1)
    t = time.perf_counter()
    for k in range(50): r = e.mads_solve(x0, 100, 1.0, seed=k)
    dt = (time.perf_counter() - t) / 50
    print(f"zerocopy={zc}: mads_solve {dt * 1e3:.3f} ms  obj {r[1]:.3f} {r[2]}")
