"""Small-swarm kernel vs CTA-per-candidate kernel on mid-size grids (which one should AUTO pick?)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
T = cov.TAN_HALF_FOV_DEFAULT
e = cov.CoverageEngine(0)
for n in (256, 512, 1024):
    bits, nf = cov.synth.fire_grid(n); d = 500 / n
    e.set_grid_bits(bits, n, n, d, d)
    for N in (5, 8):
        e.set_params(N, np.full(N, 30 * T))
        B = 200_000
        dX = e.device_alloc(B * 3 * N * 8); do = e.device_alloc(B * 8); dc = e.device_alloc(B * 8); df = e.device_alloc(B)
        e.generate_candidates(dX, B, N, seed=1)
        res = {}
        for name, k in (("auto", cov.KERNEL_AUTO), ("cta", cov.KERNEL_SPAN_GENERAL)):
            e.set_option(cov.OPT_KERNEL, k)
            for _ in range(2): e.eval_batch_device(dX, B, do, dc, df)
            e.sync(); ms0, l0 = e.kernel_time_total()
            for _ in range(3): e.eval_batch_device(dX, B, do, dc, df)
            e.sync(); ms1, l1 = e.kernel_time_total()
            res[name] = (ms1 - ms0) / (l1 - l0)
        print(f"grid {n}^2 N={N}: auto {res['auto']:.3f} ms, cta {res['cta']:.3f} ms per {B} candidates  info={e.grid_info()['planes_in_smem']}")
        for p in (dX, do, dc, df): e.device_free(p)
