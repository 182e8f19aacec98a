"""End-to-end time of cov_eval_batch (pinned host buffers) against batch size, 5 UAVs on the C2 grid: where are the
cliffs between the poll path, the single-slice path and the pipelined path?"""
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
out = {"obj": e.pinned((Bmax,)), "count": e.pinned((Bmax,), np.int64), "feasible": e.pinned((Bmax,), np.uint8)}
for B in (256, 1024, 2048, 2184, 2200, 4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288, 1048576):
    o = {k: v[:B] for k, v in out.items()}
    for _ in range(5): e.eval_batch(Xp[:B], out=o)
    reps = 200 if B <= 65536 else 20
    ms0, l0 = e.kernel_time_total()
    t = time.perf_counter()
    for _ in range(reps): e.eval_batch(Xp[:B], out=o)
    dt = (time.perf_counter() - t) / reps
    ms1, l1 = e.kernel_time_total()
    print(f"B={B:8d}: {dt * 1e6:9.1f} us per call  {B / dt / 1e6:8.1f} M evals/s   kernels {(l1 - l0) / reps:.0f}/call, {(ms1 - ms0) / reps * 1e3:8.1f} us kernel time   H2D floor {B * 120 / 55e9 * 1e6:7.1f} us")
