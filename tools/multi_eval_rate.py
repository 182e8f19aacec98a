#!/usr/bin/env python
"""cov_multi_eval_batch -- ONE process, G GPUs (what a single Julia process uses): the C2 workload, G x 1 M candidates per
call from pinned host memory, candidates sharded contiguously, host gather.  Prints evals/s for G = 1, 2, 4, 8 as far as
the box has devices, with a bit-exact check of a sample against the CPU port.
usage: python tools/multi_eval_rate.py > profiles/r2_multi_eval.json"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coverage_b200 as cov  # noqa: E402
from oracle import c_oracle  # noqa: E402

lib = cov._lib.lib
T = cov.TAN_HALF_FOV_DEFAULT
N, n = 5, 256
d = 500.0 / n
bits, n_fire = cov.synth.fire_grid(n)
r_max = np.full(N, 30 * T)
pts = cov.synth.points_from_bits(bits, n, d, d)
out = {"workload": "C2: 5 UAVs, 256x256 fire grid, 1M candidates per GPU per call, pinned host buffers", "results": []}
ndev = cov.device_count()
for G in (1, 2, 4, 8):
    if G > ndev:
        break
    devs = (C.c_int * G)(*range(G))
    m = C.c_void_p()
    assert lib.cov_multi_create(devs, G, C.byref(m)) == 0
    for k in range(G):
        h = lib.cov_multi_handle(m, k)
        assert lib.cov_set_grid_bits(h, n, n, d, d, bits.ctypes.data, d * d) == 0
        assert lib.cov_set_params(h, N, r_max.ctypes.data, 1e5, None, None, T, 0.0, 0) == 0
    B = G * 1_000_000
    eng0 = cov.CoverageEngine(0)  # only for pinned allocations
    X = eng0.pinned((B, 3 * N))
    cov.synth.random_candidates(B, N, seed=3, out=X)
    obj, cnt, fe = eng0.pinned((B,)), eng0.pinned((B,), np.int64), eng0.pinned((B,), np.uint8)
    for _ in range(3):
        assert lib.cov_multi_eval_batch(m, X.ctypes.data, B, obj.ctypes.data, cnt.ctypes.data, fe.ctypes.data) == 0
    reps = 10
    t = time.perf_counter()
    for _ in range(reps):
        assert lib.cov_multi_eval_batch(m, X.ctypes.data, B, obj.ctypes.data, cnt.ctypes.data, fe.ctypes.data) == 0
    dt = (time.perf_counter() - t) / reps
    idx = np.linspace(0, B - 1, 4000).astype(np.int64)
    want = c_oracle.eval_batch(X[idx], N, r_max, pts)
    ok = bool(np.array_equal(obj[idx], want["obj"]) and np.array_equal(cnt[idx], want["count"]))
    out["results"].append({"gpus": G, "candidates_per_call": B, "ms_per_call": dt * 1e3, "evals_per_sec": B / dt,
                           "h2d_gb_per_sec": B * 120 / dt / 1e9, "parity_on_sample": ok})
    lib.cov_multi_destroy(m)
    eng0.close()
print(json.dumps(out, indent=1))
