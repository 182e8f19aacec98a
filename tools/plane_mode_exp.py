"""CTA kernel: how the fire words are read / counted -- COV_OPT_PLANE_MODE 0 lazy, 1 early, 2 staged (TMA band),
3 paint-then-sweep -- on several shapes, with a count checksum per mode (all modes must agree) and a brute-force
re-check of the first candidates.  usage: python tools/plane_mode_exp.py [modes, e.g. 2,3]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
T = cov.TAN_HALF_FOV_DEFAULT
modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 3]
e = cov.CoverageEngine(0)
for n, N, B in ((1024, 50, 65536), (4096, 200, 8192), (1024, 13, 200000), (2048, 100, 16384), (256, 20, 100000),
                (4096, 1000, 1024), (512, 33, 50000), (1024, 9, 200000), (4096, 5, 50000), (2048, 16, 50000), (100, 5, 1000)):
    bits, nf = cov.synth.fire_grid(n, dense=(n == 100))
    d = 500 / n
    e.set_grid_bits(bits, n, n, d, d)
    e.set_params(N, np.full(N, 30 * T), sep_min=15.0)
    dX = e.device_alloc(B * 3 * N * 8)
    do, dc, df = e.device_alloc(B * 8), e.device_alloc(B * 8), e.device_alloc(B)
    e.generate_candidates(dX, B, N, seed=1)
    ref = None
    line = f"grid {n}^2 N={N} B={B}:"
    for mode in modes:
        e.set_option(cov.OPT_KERNEL, cov.KERNEL_SPAN_GENERAL)
        e.set_option(cov.OPT_PLANE_MODE, mode)
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
        line += f"  mode {mode} (ran {li['plane_mode']}, band {li['band_rows']}, grid {li['grid']}): {ms:8.3f} ms {B / ms * 1e3:.4g}/s"
        if ref is None:
            ref = cnt
        elif not np.array_equal(ref, cnt):
            line += f"  MISMATCH in {int((ref != cnt).sum())} candidates"
    nb = max(1, min(B, 4_000_000 // (nf * N // 10000 + 1) // 50))
    e.set_option(cov.OPT_KERNEL, cov.KERNEL_BRUTE)
    X = np.empty((nb, 3 * N))
    e.memcpy_d2h(X, dX)
    e.sync()
    bc = e.eval_batch(X)["count"]
    line += f"  brute[{nb}] {'ok' if np.array_equal(bc, ref[:nb]) else 'MISMATCH'}"
    print(line, flush=True)
    for p in (dX, do, dc, df):
        e.device_free(p)
