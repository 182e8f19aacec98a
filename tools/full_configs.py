"""BASELINE.json configs[2] and configs[3] at their FULL sizes on one B200, device-resident:
  C3: 50 UAVs x 4 M candidates, 1024 x 1024 fire grid, separation constraint on
  C4: 200 UAVs x 16 M candidates, 4096 x 4096 fire grid (76.8 GB of candidates in HBM), one launch
Candidates come from the library's Philox generator; a leading slice is re-evaluated with the brute-force
kernel (every cell against every disc) and compared bit for bit. Prints one JSON object per config.

usage (GPU box): python tools/full_configs.py [c3] [c4] [--out gpurun_out/full_configs.json]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import coverage_b200 as cov  # noqa: E402

T = cov.TAN_HALF_FOV_DEFAULT
CONFIGS = {"c3": (1024, 50, 4_000_000, 2048), "c4": (4096, 200, 16_000_000, 512)}


def run(e, name):
    n, N, B, n_check = CONFIGS[name]
    bits, n_fire = cov.synth.fire_grid(n)
    d = 500.0 / n
    e.set_grid_bits(bits, n, n, d, d)
    e.set_params(N, np.full(N, 30 * T), sep_min=15.0)
    dX = e.device_alloc(B * 3 * N * 8)
    do, dc, df = e.device_alloc(B * 8), e.device_alloc(B * 8), e.device_alloc(B)
    e.generate_candidates(dX, B, N, seed=7)
    e.sync()
    e.set_option(cov.OPT_KERNEL, cov.KERNEL_AUTO)
    e.eval_batch_device(dX, min(B, 65536), do, dc, df)  # warm-up (module load, smem carve-out)
    e.sync()
    ms0, l0 = e.kernel_time_total()
    t0 = time.perf_counter()
    e.eval_batch_device(dX, B, do, dc, df)
    e.sync()
    wall = time.perf_counter() - t0
    ms1, l1 = e.kernel_time_total()
    ll = e.last_launch()
    count = np.empty(B, np.int64)
    obj = np.empty(B, np.float64)
    feas = np.empty(B, np.uint8)
    e.memcpy_d2h(count, dc); e.memcpy_d2h(obj, do); e.memcpy_d2h(feas, df)
    e.sync()
    # the first and the last candidates (byte offsets beyond 2^32 on C4) again through the brute-force kernel
    e.set_option(cov.OPT_KERNEL, cov.KERNEL_BRUTE)
    ok = True
    for first in (0, B - n_check):
        e.eval_batch_device(dX + first * 3 * N * 8, n_check, do, dc, df)
        e.sync()
        c2, o2, f2 = np.empty(n_check, np.int64), np.empty(n_check, np.float64), np.empty(n_check, np.uint8)
        e.memcpy_d2h(c2, dc); e.memcpy_d2h(o2, do); e.memcpy_d2h(f2, df)
        e.sync()
        sl = slice(first, first + n_check)
        ok = ok and bool(np.array_equal(c2, count[sl]) and np.array_equal(o2.view(np.uint64), obj[sl].view(np.uint64))
                         and np.array_equal(f2, feas[sl]))
    e.set_option(cov.OPT_KERNEL, cov.KERNEL_AUTO)
    for p in (dX, do, dc, df):
        e.device_free(p)
    kernel_s = (ms1 - ms0) / 1e3
    return {"config": name, "grid": f"{n}x{n}", "fire_cells": int(n_fire), "uavs": N, "candidates": B,
            "candidate_bytes_in_hbm": B * 3 * N * 8, "launches": l1 - l0, "kernel_s": kernel_s, "wall_s": wall,
            "evals_per_s": B / kernel_s, "algorithmic_tests_per_s": B * float(n_fire) * N / kernel_s,
            "kernel": {1: "span_small", 4: "span_cta"}.get(ll["kernel"], ll["kernel"]), "launch": ll,
            "count_sum": int(count.sum()), "count_max": int(count.max()), "feasible": int(feas.sum()),
            "brute_check": {"candidates": 2 * n_check, "where": "first and last", "bit_exact": ok}}


def main():
    names = [a for a in sys.argv[1:] if a in CONFIGS] or ["c3", "c4"]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    res = []
    with cov.CoverageEngine(0) as e:
        for nm in names:
            r = run(e, nm)
            print(json.dumps(r), flush=True)
            res.append(r)
    if out:
        json.dump(res, open(out, "w"), indent=1)
    assert all(r["brute_check"]["bit_exact"] for r in res)


if __name__ == "__main__":
    main()
