import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
T = cov.TAN_HALF_FOV_DEFAULT
e = cov.CoverageEngine(0)
for n, N, B in ((1024, 50, 32768), (4096, 200, 4096)):
    bits, nf = cov.synth.fire_grid(n); d = 500 / n
    e.set_grid_bits(bits, n, n, d, d)
    e.set_params(N, np.full(N, 30 * T), sep_min=15.0)
    dX = e.device_alloc(B * 3 * N * 8); do = e.device_alloc(B * 8); dc = e.device_alloc(B * 8); df = e.device_alloc(B)
    e.generate_candidates(dX, B, N, seed=1)
    for ctas in (1, 2, 3):
        e.set_option(cov.OPT_CTAS_PER_SM, ctas)
        for _ in range(2): e.eval_batch_device(dX, B, do, dc, df)
        e.sync(); ms0, l0 = e.kernel_time_total()
        for _ in range(3): e.eval_batch_device(dX, B, do, dc, df)
        e.sync(); ms1, l1 = e.kernel_time_total()
        print(f"grid {n}^2 N={N} B={B} CTAs/SM {ctas}: {(ms1 - ms0) / (l1 - l0):8.3f} ms")
    e.set_option(cov.OPT_CTAS_PER_SM, 0)
    for p in (dX, do, dc, df): e.device_free(p)
