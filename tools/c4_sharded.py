"""BASELINE.json configs[3] at full size, candidate-sharded over 1 / 2 / 4 / 8 B200 of one box (strong scaling):
200 UAVs x 16 M candidates on the 4096 x 4096 fire grid.  One process, one engine (handle) per device - the
pattern cov_multi_* wraps; shard r of G holds candidates [r*B/G, (r+1)*B/G) of the SAME Philox sequence, so the
sum of all counts must be identical for every G (and to tools/full_configs.py's single-GPU run).  No data-path
collective: only the per-shard results are gathered.  Time = slowest shard's kernel time (CUDA events).

usage (8-GPU box): python tools/c4_sharded.py [--out gpurun_out/c4_sharded.json] [--candidates 16000000]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import coverage_b200 as cov  # noqa: E402

T = cov.TAN_HALF_FOV_DEFAULT
n, N = 4096, 200
B = int(sys.argv[sys.argv.index("--candidates") + 1]) if "--candidates" in sys.argv else 16_000_000
out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None

n_dev = cov.device_count()
bits, n_fire = cov.synth.fire_grid(n)
d = 500.0 / n
engines = []
for k in range(n_dev):
    e = cov.CoverageEngine(k)
    e.set_grid_bits(bits, n, n, d, d)
    e.set_params(N, np.full(N, 30 * T), sep_min=15.0)
    engines.append(e)

runs = []
for G in (1, 2, 4, 8):
    if G > n_dev:
        break
    shard = B // G
    bufs = []
    for r in range(G):
        e = engines[r]
        dX = e.device_alloc(shard * 3 * N * 8)
        do, dc, df = e.device_alloc(shard * 8), e.device_alloc(shard * 8), e.device_alloc(shard)
        e.generate_candidates(dX, shard, N, seed=7, first_index=r * shard)
        e.eval_batch_device(dX, 4096, do, dc, df)  # warm-up
        bufs.append((dX, do, dc, df))
    for r in range(G):
        engines[r].sync()
    before = [engines[r].kernel_time_total() for r in range(G)]
    t0 = time.perf_counter()
    for r in range(G):
        engines[r].eval_batch_device(bufs[r][0], shard, *bufs[r][1:])
    for r in range(G):
        engines[r].sync()
    wall = time.perf_counter() - t0
    after = [engines[r].kernel_time_total() for r in range(G)]
    dev_s = [(a[0] - b[0]) / 1e3 for a, b in zip(after, before)]
    count_sum, best = 0, (np.inf, -1)
    for r in range(G):
        e = engines[r]
        cnt, obj = np.empty(shard, np.int64), np.empty(shard, np.float64)
        e.memcpy_d2h(cnt, bufs[r][2]); e.memcpy_d2h(obj, bufs[r][1])
        e.sync()
        count_sum += int(cnt.sum())
        i = int(np.argmin(obj))
        best = min(best, (float(obj[i]), r * shard + i))
        for p in bufs[r]:
            e.device_free(p)
    t = max(dev_s)
    runs.append({"n_gpus": G, "candidates": shard * G, "per_gpu": shard, "kernel_s_max": t, "kernel_s_per_gpu": dev_s,
                 "wall_s": wall, "evals_per_s": shard * G / t, "algorithmic_tests_per_s": shard * G * float(n_fire) * N / t,
                 "count_sum": count_sum, "argmin": {"obj": best[0], "index": best[1]}})
    print(json.dumps(runs[-1]), flush=True)

assert len({r["count_sum"] for r in runs}) == 1 and len({r["argmin"]["index"] for r in runs}) == 1
res = {"workload": f"C4 full size: {N} UAVs x {B} candidates, {n}x{n} fire grid ({n_fire} burning cells), strong scaling",
       "runs": runs, "speedup_vs_1": [runs[0]["kernel_s_max"] / r["kernel_s_max"] for r in runs]}
print(json.dumps(res["speedup_vs_1"]))
if out:
    json.dump(res, open(out, "w"), indent=1)
