"""Latency of poll-sized calls (what DirectSearch.jl / a MADS poll driver issue): the scalar closure, a
30-point poll set, 128 candidates; zero-copy path (default) vs the copy-engine path (COV_OPT_ZEROCOPY_OUT=0)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
T = cov.TAN_HALF_FOV_DEFAULT
for label, setup in (("C1 grid 100x100", lambda: e.set_grid_full(100, 100, 5.0, 5.0)),
                     ("C2 grid 256x256", lambda: e.set_grid_bits(cov.synth.fire_grid(256)[0], 256, 256, 500 / 256, 500 / 256))):
    setup()
    e.set_params(5, np.full(5, 30 * T))
    X = cov.synth.random_candidates(4096, 5, seed=3)
    for zc in (1, 0):
        e.set_option(cov.OPT_ZEROCOPY_OUT, zc)
        row = []
        x = X[0].copy()
        for _ in range(20): e.eval_one(x)
        t = time.perf_counter()
        for _ in range(2000): e.eval_one(x)
        row.append("eval_one %.1f us" % ((time.perf_counter() - t) / 2000 * 1e6))
        for nb in (30, 128, 200):
            P = X[:nb].copy()
            out = e.eval_batch(P)
            for _ in range(20): e.eval_batch(P, out=out)
            ms0, l0 = e.kernel_time_total()
            t = time.perf_counter()
            for _ in range(2000): e.eval_batch(P, out=out)
            dt = time.perf_counter() - t
            ms1, l1 = e.kernel_time_total()
            row.append("batch(%d) %.1f us (kernel %.1f us)" % (nb, dt / 2000 * 1e6, (ms1 - ms0) / (l1 - l0) * 1e3))
        ref = e.eval_batch(X[:200].copy())
        print(f"{label}  zerocopy={zc}: " + ", ".join(row) + f"  checksum {int(ref['count'].sum())}")
    e.set_option(cov.OPT_ZEROCOPY_OUT, 1)
