import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
T = cov.TAN_HALF_FOV_DEFAULT
e.set_grid_bits(cov.synth.fire_grid(256)[0], 256, 256, 500 / 256, 500 / 256)
N = 5
e.set_params(N, np.full(N, 30 * T))
Bmax = 1 << 18
X = cov.synth.random_candidates(Bmax, N, seed=3)
Xp = e.pinned((Bmax, 3 * N)); Xp[:] = X
outp = {"obj": e.pinned((Bmax,)), "count": e.pinned((Bmax,), np.int64), "feasible": e.pinned((Bmax,), np.uint8)}
outg = {"obj": np.empty(Bmax), "count": np.empty(Bmax, np.int64), "feasible": np.empty(Bmax, np.uint8)}
for B in (16384, 65536, 262144):
    for iname, xin in (("in pageable", X), ("in pinned", Xp)):
        for oname, out in (("out pageable", outg), ("out pinned", outp)):
            o = {k: v[:B] for k, v in out.items()}
            for _ in range(3): e.eval_batch(xin[:B], out=o)
            t = time.perf_counter()
            for _ in range(50): e.eval_batch(xin[:B], out=o)
            print(f"B={B:7d} {iname:12s} {oname:13s} {(time.perf_counter() - t) / 50 * 1e6:9.1f} us")
    t = time.perf_counter()
    for _ in range(50): Xp[:B] = X[:B]
    print(f"   numpy copy of the input: {(time.perf_counter() - t) / 50 * 1e6:9.1f} us")
