"""cov_eval_batch on PAGEABLE host buffers (what a Julia caller passes) against batch size; pinned beside it."""
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
Xp = e.pinned((Bmax, 3 * N)); Xp[:] = X
outp = {"obj": e.pinned((Bmax,)), "count": e.pinned((Bmax,), np.int64), "feasible": e.pinned((Bmax,), np.uint8)}
outg = {"obj": np.empty(Bmax), "count": np.empty(Bmax, np.int64), "feasible": np.empty(Bmax, np.uint8)}
for B in (2200, 4096, 16384, 65536, 131072, 262144, 524288, 1048576):
    row = []
    for name, xin, out in (("pageable", X, outg), ("pinned", Xp, outp)):
        o = {k: v[:B] for k, v in out.items()}
        for _ in range(3): e.eval_batch(xin[:B], out=o)
        reps = 100 if B <= 65536 else 20
        t = time.perf_counter()
        for _ in range(reps): e.eval_batch(xin[:B], out=o)
        row.append("%s %9.1f us" % (name, (time.perf_counter() - t) / reps * 1e6))
    print(f"B={B:8d}: " + "   ".join(row))
