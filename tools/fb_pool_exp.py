"""(Needs the COV_OPT_FB_POOL experiment build; the option is not in the shipped library — kept for the record.)
Small-swarm kernel: one framebuffer per warp (20 warps/SM at 256 columns) vs a pool shared by 24 warps
(COV_OPT_FB_POOL). Bench distribution (about a quarter of the candidates have overlapping discs) and a
clustered one (every candidate overlaps: the pool's worst case)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
T = cov.TAN_HALF_FOV_DEFAULT
bits, nf = cov.synth.fire_grid(256); d = 500 / 256
e.set_grid_bits(bits, 256, 256, d, d)
B = 1_000_000
for N in (5, 8):
    e.set_params(N, np.full(N, 30 * T))
    dX = e.device_alloc(B * 3 * N * 8); do = e.device_alloc(B * 8); dc = e.device_alloc(B * 8); df = e.device_alloc(B)
    for dist in ("bench", "clustered"):
        if dist == "bench":
            e.generate_candidates(dX, B, N, seed=1)
        else:
            rng = np.random.default_rng(0)
            X = np.concatenate([200 + rng.random((B, 2 * N)) * 60, (15 + rng.random((B, N)) * 15) * T], axis=1)
            e.memcpy_h2d(dX, X); e.sync()
        for pool in (0, 1, 0, 1):
            e.set_option(cov.OPT_FB_POOL, pool)
            for _ in range(3): e.eval_batch_device(dX, B, do, dc, df)
            e.sync(); ms0, l0 = e.kernel_time_total()
            for _ in range(10): e.eval_batch_device(dX, B, do, dc, df)
            e.sync(); ms1, l1 = e.kernel_time_total()
            cnt = np.empty(B, np.int64); e.memcpy_d2h(cnt, dc); e.sync()
            ll = e.last_launch()
            print(f"N={N} {dist:9s} pool={pool}: {(ms1 - ms0) / (l1 - l0):7.4f} ms  block {ll['block']} smem {ll['smem_bytes']}  checksum {int(cnt.sum())}")
    for p in (dX, do, dc, df): e.device_free(p)
e.set_option(cov.OPT_FB_POOL, 0)
