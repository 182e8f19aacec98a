"""Which kernel should take a small batch? Device time of the small-swarm kernel (warp per 8..32 candidates) vs the
CTA-per-candidate kernel for B = 32 .. 4096 on the C1 and C2 grids (device-resident candidates)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
T = cov.TAN_HALF_FOV_DEFAULT
for label, setup in (("C1 100x100", lambda: e.set_grid_full(100, 100, 5.0, 5.0)),
                     ("C2 256x256", lambda: e.set_grid_bits(cov.synth.fire_grid(256)[0], 256, 256, 500 / 256, 500 / 256))):
    setup()
    for N in (5, 8):
        e.set_params(N, np.full(N, 30 * T))
        Bmax = 4096
        dX = e.device_alloc(Bmax * 3 * N * 8); do = e.device_alloc(Bmax * 8); dc = e.device_alloc(Bmax * 8); df = e.device_alloc(Bmax)
        e.generate_candidates(dX, Bmax, N, seed=1)
        for B in (32, 64, 128, 200, 300, 444, 600, 888, 1332, 2048, 4096):
            res = []
            for name, k in (("small", cov.KERNEL_SPAN), ("cta", cov.KERNEL_SPAN_GENERAL)):
                e.set_option(cov.OPT_KERNEL, k)
                for _ in range(5): e.eval_batch_device(dX, B, do, dc, df)
                e.sync(); ms0, l0 = e.kernel_time_total()
                for _ in range(50): e.eval_batch_device(dX, B, do, dc, df)
                e.sync(); ms1, l1 = e.kernel_time_total()
                res.append("%s[%d] %.1f us" % (name, e.last_launch()["kernel"], (ms1 - ms0) / (l1 - l0) * 1e3))
            print(f"{label} N={N} B={B}: " + "  ".join(res))
        e.set_option(cov.OPT_KERNEL, cov.KERNEL_AUTO)
        for p in (dX, do, dc, df): e.device_free(p)
