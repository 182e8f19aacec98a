"""Slice size of the host pipeline for PAGEABLE buffers (staging memcpy + DMA + kernel overlap), 5 UAVs, C2 grid."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
T = cov.TAN_HALF_FOV_DEFAULT
e.set_grid_bits(cov.synth.fire_grid(256)[0], 256, 256, 500 / 256, 500 / 256)
N = 5
e.set_params(N, np.full(N, 30 * T))
Bmax = 1 << 20
X = cov.synth.random_candidates(Bmax, N, seed=3)
out = {"obj": np.empty(Bmax), "count": np.empty(Bmax, np.int64), "feasible": np.empty(Bmax, np.uint8)}
for B in (32768, 65536, 131072, 262144, 524288, 1048576):
    row = []
    for c in (0, 16384, 32768, 65536):
        if c >= B:
            continue
        e.set_option(cov.OPT_CHUNK, c)
        o = {k: v[:B] for k, v in out.items()}
        for _ in range(3): e.eval_batch(X[:B], out=o)
        reps = 50 if B <= 131072 else 15
        t = time.perf_counter()
        for _ in range(reps): e.eval_batch(X[:B], out=o)
        row.append("%s: %8.1f" % (c if c else "auto", (time.perf_counter() - t) / reps * 1e6))
    print(f"B={B:8d} us/call  " + "   ".join(row))
e.set_option(cov.OPT_CHUNK, 0)
