"""CTA kernel, wide grids: how many co-resident CTAs vs how many rows per band (COV_MIN_BAND_ROWS in cta_plan).
Run once per variant build: COVERAGE_CUDA_LIB=build/variants/lib_tNN.so python tools/band_threshold_exp.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
T = cov.TAN_HALF_FOV_DEFAULT
e = cov.CoverageEngine(0)
for n, N, B in ((4096, 5, 50000), (4096, 20, 30000), (4096, 50, 20000), (4096, 100, 12000), (4096, 200, 8192), (4096, 400, 3000),
                (4096, 1000, 1024), (8192, 50, 8000), (8192, 200, 3000), (2048, 400, 4000), (2048, 1000, 1500), (1024, 1000, 3000)):
    bits, nf = cov.synth.fire_grid(n)
    d = 500 / n
    e.set_grid_bits(bits, n, n, d, d)
    e.set_params(N, np.full(N, 30 * T), sep_min=15.0)
    dX = e.device_alloc(B * 3 * N * 8)
    do, dc, df = e.device_alloc(B * 8), e.device_alloc(B * 8), e.device_alloc(B)
    e.generate_candidates(dX, B, N, seed=1)
    e.set_option(cov.OPT_KERNEL, cov.KERNEL_SPAN_GENERAL)
    for _ in range(2):
        e.eval_batch_device(dX, B, do, dc, df)
    e.sync()
    ms0, l0 = e.kernel_time_total()
    for _ in range(3):
        e.eval_batch_device(dX, B, do, dc, df)
    e.sync()
    ms1, l1 = e.kernel_time_total()
    cnt = np.empty(B, np.int64)
    e.memcpy_d2h(cnt, dc)
    e.sync()
    li = e.last_launch()
    ms = (ms1 - ms0) / (l1 - l0)
    print(f"grid {n}^2 N={N} B={B}: mode {li['plane_mode']} band {li['band_rows']} grid {li['grid']} block {li['block']}: "
          f"{ms:8.3f} ms  checksum {int(cnt.sum())}", flush=True)
    for p in (dX, do, dc, df):
        e.device_free(p)
