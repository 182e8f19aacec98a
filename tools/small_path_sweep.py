import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
T = cov.TAN_HALF_FOV_DEFAULT
e.set_grid_bits(cov.synth.fire_grid(256)[0], 256, 256, 500 / 256, 500 / 256)
N = 5
e.set_params(N, np.full(N, 30 * T))
X = cov.synth.random_candidates(4096, N, seed=3)
for B in (30, 128, 256, 512, 1024, 1500, 2184):
    P = X[:B].copy()
    out = e.eval_batch(P)
    for _ in range(20): e.eval_batch(P, out=out)
    ms0, l0 = e.kernel_time_total()
    t = time.perf_counter()
    for _ in range(1000): e.eval_batch(P, out=out)
    dt = (time.perf_counter() - t) / 1000
    ms1, l1 = e.kernel_time_total()
    print(f"B={B:5d}: {dt * 1e6:7.1f} us per call (kernel {(ms1 - ms0) / (l1 - l0) * 1e3:5.1f} us) checksum {int(out['count'].sum())}")
