"""Slice size of the host pipeline against batch size (pinned buffers, 5 UAVs, C2 grid): ms per call."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
bits, n = cov.synth.fire_grid(256); d = 500 / 256
e.set_grid_bits(bits, 256, 256, d, d); e.set_params(5, np.full(5, 30 * cov.TAN_HALF_FOV_DEFAULT))
Bmax = 1 << 20
X = e.pinned((Bmax, 15)); cov.synth.random_candidates(Bmax, 5, seed=3, out=X)
out = {"obj": e.pinned((Bmax,)), "count": e.pinned((Bmax,), np.int64), "feasible": e.pinned((Bmax,), np.uint8)}
for B in (16384, 32768, 65536, 131072, 262144, 524288, 1000000):
    row = []
    o = {k: v[:B] for k, v in out.items()}
    for slice_c in (0, 8192, 16384, 32768, 65536, 139808):
        if slice_c and slice_c >= B and slice_c != 139808:
            continue
        e.set_option(cov.OPT_CHUNK, slice_c)
        for _ in range(3): e.eval_batch(X[:B], out=o)
        reps = 100 if B <= 131072 else 30
        t = time.perf_counter()
        for _ in range(reps): e.eval_batch(X[:B], out=o)
        row.append("%s: %.1f" % (slice_c if slice_c else "auto", (time.perf_counter() - t) / reps * 1e6))
    print(f"B={B:8d} us/call  " + "  ".join(row))
e.set_option(cov.OPT_CHUNK, 0)
