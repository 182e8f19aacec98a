"""End-to-end time of cov_eval_batch_packed (pinned host buffers, 5 UAVs on the C2 grid, 1M candidates per call) by
element type and slice size (COV_OPT_CHUNK), beside cov_eval_batch on the widened Float64 matrix."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
T = cov.TAN_HALF_FOV_DEFAULT
e.set_grid_bits(cov.synth.fire_grid(256)[0], 256, 256, 500 / 256, 500 / 256)
N, B, SETS = 5, 1_000_000, 4
e.set_params(N, np.full(N, 30 * T))
out = {"obj": e.pinned((B,)), "count": e.pinned((B,), np.int64), "feasible": e.pinned((B,), np.uint8)}


def timed(fn, reps=20):
    for k in range(4): fn(k)
    ms0, l0 = e.kernel_time_total()
    t = time.perf_counter()
    for k in range(reps): fn(k)
    dt = (time.perf_counter() - t) / reps
    ms1, l1 = e.kernel_time_total()
    return dt * 1e3, (ms1 - ms0) / reps


Q16 = [e.pinned((B, 3 * N), np.int16) for _ in range(SETS)]
for k in range(SETS): cov.synth.mesh_candidates(B, N, seed=k, out=Q16[k])
X = [e.pinned((B, 3 * N)) for _ in range(SETS)]
for k in range(SETS): X[k][:] = Q16[k]
ms, kms = timed(lambda k: e.eval_batch(X[k % SETS], out=out))
ref = {k: np.array(v) for k, v in e.eval_batch(X[0], out=out).items()}
print(f"float64 (cov_eval_batch)     : {ms:6.3f} ms per call  {B / ms / 1e3:7.1f} M evals/s  coverage kernels {kms:5.3f} ms  [H2D floor {B * 120 / 55.4e9 * 1e3:5.3f} ms]")
for name, dt in (("int16", np.int16), ("int32", np.int32), ("float32", np.float32)):
    Q = [e.pinned((B, 3 * N), dt) for _ in range(SETS)]
    for k in range(SETS): Q[k][:] = Q16[k]
    for chunk in (0, 32768, 65536, 262144, 524288, 1 << 20):
        e.set_option(cov.OPT_CHUNK, chunk)
        ms, kms = timed(lambda k: e.eval_batch_packed(Q[k % SETS], 1.0, out=out))
        r = e.eval_batch_packed(Q[0], 1.0, out=out)
        same = all(np.array_equal(r[k], ref[k]) for k in ref)
        print(f"{name:8s} slice {chunk or 'auto':>8}: {ms:6.3f} ms per call  {B / ms / 1e3:7.1f} M evals/s  coverage kernels {kms:5.3f} ms  "
              f"[H2D floor {B * 15 * np.dtype(dt).itemsize / 55.4e9 * 1e3:5.3f} ms]  same={same}")
    e.set_option(cov.OPT_CHUNK, 0)
# where the time goes: the per-slice device timeline of one int16 call (COV_OPT_TRACE)
e.set_option(cov.OPT_TRACE, 1)
e.eval_batch_packed(Q16[0], 1.0, out=out)
t = e.trace()
e.set_option(cov.OPT_TRACE, 0)
print("int16 timeline, ms since the first copy was queued [h2d done, kernels start, kernels end, results done] per slice:")
for row in t: print("   ", " ".join(f"{v:7.3f}" for v in row))
# pageable int16 (a plain Julia Array{Int16})
Qp = [np.array(q) for q in Q16]
outp = {"obj": np.empty(B), "count": np.empty(B, dtype=np.int64), "feasible": np.empty(B, dtype=np.uint8)}
ms, kms = timed(lambda k: e.eval_batch_packed(Qp[k % SETS], 1.0, out=outp))
print(f"int16 pageable in/out        : {ms:6.3f} ms per call  {B / ms / 1e3:7.1f} M evals/s")
