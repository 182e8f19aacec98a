"""Throughput of the continuous union-area variant (cov_union_area_batch, host buffers): 5 / 50 / 200 discs."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
rng = np.random.default_rng(0)
for N, B in ((5, 1_000_000), (50, 100_000), (64, 100_000), (65, 20_000), (200, 20_000), (1024, 1_000)):
    X = np.concatenate([rng.random((B, 2 * N)) * 500, (5 + rng.random((B, N)) * 25) * cov.TAN_HALF_FOV_DEFAULT], axis=1)
    e.union_area(X[:1000], N)
    t = time.perf_counter()
    a = e.union_area(X, N)
    dt = time.perf_counter() - t
    print(f"N={N:5d} B={B:8d}: {dt * 1e3:9.2f} ms  {B / dt:12.4g} areas/s  mean area {a.mean():.1f}")
