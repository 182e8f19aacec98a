"""Small-swarm kernel A/B: C2 and C1 shapes, 1 M device-resident candidates, event-timed kernel ms (mean of 10 launches
over 4 rotating sets), a count checksum (must not change between library variants) and a brute-force re-check of
20 000 candidates.  usage: COVERAGE_CUDA_LIB=build/variants/lib_x.so python tools/c2_quick.py"""
import os
import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
T = cov.TAN_HALF_FOV_DEFAULT
e = cov.CoverageEngine(0)
tag = os.path.basename(os.environ.get("COVERAGE_CUDA_LIB", "main"))
for name, n, dense, N, sep in (("C2", 256, False, 5, 0.0), ("C1", 100, True, 5, 0.0), ("C2x8", 256, False, 8, 15.0)):
    bits, nf = cov.synth.fire_grid(n, dense=dense)
    d = 5.0 if n == 100 else 500 / n
    e.set_grid_bits(bits, n, n, d, d)
    e.set_params(N, np.full(N, 30 * T), sep_min=sep)
    e.set_option(cov.OPT_KERNEL, cov.KERNEL_AUTO)
    B = 1_000_000
    sets = [e.device_alloc(B * 3 * N * 8) for _ in range(4)]
    do, dc, df = e.device_alloc(B * 8), e.device_alloc(B * 8), e.device_alloc(B)
    for k, p in enumerate(sets):
        e.generate_candidates(p, B, N, seed=1, first_index=k * B)
    for k in range(3):
        e.eval_batch_device(sets[k % 4], B, do, dc, df)
    e.sync()
    ms0, l0 = e.kernel_time_total()
    for k in range(12):
        e.eval_batch_device(sets[k % 4], B, do, dc, df)
    e.sync()
    ms1, l1 = e.kernel_time_total()
    cnt = np.empty(B, np.int64)
    e.memcpy_d2h(cnt, dc)
    X = np.empty((20000, 3 * N))
    e.memcpy_d2h(X, sets[3])
    e.sync()
    li = e.last_launch()
    e.set_option(cov.OPT_KERNEL, cov.KERNEL_BRUTE)
    bc = e.eval_batch(X)["count"]
    ok = np.array_equal(bc, cnt[:20000])
    ms = (ms1 - ms0) / (l1 - l0)
    print(f"{tag:18s} {name:5s} chunk {li['chunk']:2d} warps {li['block'] // 32:2d}: {ms:7.4f} ms  {B / ms * 1e3:.4g} evals/s  "
          f"checksum {int(cnt.sum())}  brute[20000] {'ok' if ok else 'MISMATCH'}", flush=True)
    for p in sets + [do, dc, df]:
        e.device_free(p)
