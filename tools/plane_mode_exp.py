"""CTA kernel: fire words lazily / early through L2, or staged per band with TMA (which is fastest where?)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
T = cov.TAN_HALF_FOV_DEFAULT
e = cov.CoverageEngine(0)
for n, N, B in ((1024, 50, 32768), (4096, 200, 4096), (1024, 5, 200000), (2048, 20, 65536)):
    bits, nf = cov.synth.fire_grid(n); d = 500 / n
    e.set_grid_bits(bits, n, n, d, d)
    e.set_params(N, np.full(N, 30 * T), sep_min=15.0 if N >= 50 else 0.0)
    dX = e.device_alloc(B * 3 * N * 8); do = e.device_alloc(B * 8); dc = e.device_alloc(B * 8); df = e.device_alloc(B)
    e.generate_candidates(dX, B, N, seed=1)
    ref = None
    for mode in (0, 1, 2):
        e.set_option(cov.OPT_PLANE_MODE, mode)
        for _ in range(2): e.eval_batch_device(dX, B, do, dc, df)
        e.sync(); ms0, l0 = e.kernel_time_total()
        for _ in range(3): e.eval_batch_device(dX, B, do, dc, df)
        e.sync(); ms1, l1 = e.kernel_time_total()
        c = np.empty(B, dtype=np.int64); e.memcpy_d2h(c, dc); e.sync()
        if ref is None: ref = c.copy()
        print(f"grid {n}^2 N={N} B={B} mode {mode}: {(ms1 - ms0) / (l1 - l0):8.3f} ms  same counts: {np.array_equal(c, ref)}")
    e.set_option(cov.OPT_PLANE_MODE, -1)
    for p in (dX, do, dc, df): e.device_free(p)
