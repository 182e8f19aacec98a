import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
T = cov.TAN_HALF_FOV_DEFAULT
e = cov.CoverageEngine(0)
for n, N, B in ((1024, 50, 65536), (4096, 200, 8192), (1024, 13, 200000), (2048, 100, 16384)):
    bits, nf = cov.synth.fire_grid(n); d = 500 / n
    e.set_grid_bits(bits, n, n, d, d)
    e.set_params(N, np.full(N, 30 * T), sep_min=15.0)
    dX = e.device_alloc(B * 3 * N * 8); do = e.device_alloc(B * 8); dc = e.device_alloc(B * 8); df = e.device_alloc(B)
    e.generate_candidates(dX, B, N, seed=1)
    for _ in range(2): e.eval_batch_device(dX, B, do, dc, df)
    e.sync(); ms0, l0 = e.kernel_time_total()
    for _ in range(3): e.eval_batch_device(dX, B, do, dc, df)
    e.sync(); ms1, l1 = e.kernel_time_total()
    cnt = np.empty(B, np.int64); e.memcpy_d2h(cnt, dc); e.sync()
    print(f"grid {n}^2 N={N} B={B}: {(ms1 - ms0) / (l1 - l0):9.3f} ms  {B / ((ms1 - ms0) / (l1 - l0)) * 1e3:.4g} evals/s  checksum {int(cnt.sum())}")
    for p in (dX, do, dc, df): e.device_free(p)
