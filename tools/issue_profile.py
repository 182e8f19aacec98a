#!/usr/bin/env python
"""Executed warp instructions per candidate of every benchmark kernel -> profiles/r2_issue.json
(the numerator of bench.py's issue-slot roofline).

On the GPU box (one ncu pass, a handful of metrics, every coverage-kernel launch of the run):

    python tools/issue_profile.py run > gpurun_out/issue_plain.log 2>&1 &&
    ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:'span_small_kernel|span_cta_kernel' --csv \
        --log-file gpurun_out/issue.csv python tools/issue_profile.py run

`run` launches, for each bench workload (bench.py WORKLOADS: c2, c3, c4, c1), the device-resident batch the bench
times -- one warm-up launch and two more -- and writes gpurun_out/issue_run.json: the ordered list of coverage-kernel
launches with (workload, kernel instantiation, candidates).  Here, on the CPU box:

    python tools/issue_profile.py parse gpurun_out/issue.csv gpurun_out/issue_run.json

matches the i-th profiled coverage kernel with the i-th recorded launch and writes profiles/r2_issue.json, stamped
with the hash of csrc/ (bench.py flags a stale profile).
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import numpy as np
    import bench
    import coverage_b200 as cov
    launches = []
    for name in ("c2", "c3", "c4", "c1"):
        wl = bench.WORKLOADS[name]
        N, B = wl["n"], wl["batch"]
        bits, n_fire, d, r_max = bench.make_workload(cov.synth, wl)
        eng = cov.CoverageEngine(0)
        eng.set_grid_bits(bits, wl["grid"], wl["grid"], d, d)
        eng.set_params(N, r_max, sep_min=wl["sep"])
        dX = eng.device_alloc(B * 3 * N * 8)
        d_obj, d_cnt, d_fe = eng.device_alloc(B * 8), eng.device_alloc(B * 8), eng.device_alloc(B)
        for k in range(3):
            eng.generate_candidates(dX, B, N, seed=1, first_index=k * B)
            eng.eval_batch_device(dX, B, d_obj, d_cnt, d_fe)
            eng.sync()
            launches.append({"workload": name, "kernel": bench.kernel_key(eng.last_launch()), "candidates": B,
                             "warm": k == 0})
        cnt = np.empty(B, dtype=np.int64)
        eng.memcpy_d2h(cnt, d_cnt)
        eng.sync()
        launches[-1]["count_sum"] = int(cnt.sum())
        eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "issue_run.json"), "w") as f:
        json.dump(launches, f, indent=1)
    print(json.dumps(launches))


def parse(csv_path, run_path):
    import bench
    with open(run_path) as f:
        launches = json.load(f)
    with open(csv_path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    # long format: one row per (launch ID, metric)
    by_id = {}
    order = []
    for r in rows:
        kid = r["ID"]
        if kid not in by_id:
            by_id[kid] = {"name": r["Kernel Name"]}
            order.append(kid)
        try:
            by_id[kid][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
        by_id[kid][r["Metric Name"] + "|unit"] = r["Metric Unit"]
    prof = [by_id[k] for k in order if "span_small_kernel" in by_id[k]["name"] or "span_cta_kernel" in by_id[k]["name"]]
    if len(prof) != len(launches):
        raise SystemExit(f"{len(prof)} profiled coverage kernels but {len(launches)} recorded launches")
    kernels = {}
    for p, l in zip(prof, launches):
        if l["warm"]:
            continue
        key = f"{l['workload']}|{l['kernel']}"
        base = l["kernel"].split("<")[0]
        assert base in p["name"], (p["name"], l["kernel"])
        dur = p["gpu__time_duration.sum"]
        unit = p.get("gpu__time_duration.sum|unit", "ns")
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        ent = kernels.setdefault(key, {"samples": 0, "warp_instr": 0.0, "dram": 0.0, "issue": 0.0, "warps": 0.0, "ms": 0.0})
        ent["samples"] += 1
        ent["warp_instr"] += p["smsp__inst_executed.sum"]
        ent["dram"] += p.get("dram__bytes_read.sum", 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(
            p.get("dram__bytes_read.sum|unit", "byte"), 1) + p.get("dram__bytes_write.sum", 0.0) * {
            "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(p.get("dram__bytes_write.sum|unit", "byte"), 1)
        ent["issue"] += p.get("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        ent["warps"] += p.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
        ent["ms"] += dur_ms
        ent["candidates"] = l["candidates"]
        ent["ncu_kernel_name"] = p["name"]
    out = {"source_sha": bench.source_sha(),
           "how": "ncu --metrics smsp__inst_executed.sum,... --clock-control none on tools/issue_profile.py run "
                  "(device-resident Philox candidates, the bench workloads at their bench batch sizes; mean of 2 launches)",
           "kernels": {}}
    for key, e in sorted(kernels.items()):
        n = e["samples"]
        out["kernels"][key] = {
            "warp_instr_per_candidate": e["warp_instr"] / n / e["candidates"],
            "warp_instr_per_launch": e["warp_instr"] / n, "candidates_per_launch": e["candidates"],
            "dram_bytes_per_launch": e["dram"] / n, "issue_active_pct": e["issue"] / n,
            "warps_active_pct": e["warps"] / n, "ncu_kernel_ms": e["ms"] / n, "ncu_kernel_name": e["ncu_kernel_name"]}
    dst = os.path.join(ROOT, "profiles", "r2_issue.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "run":
        run()
    elif len(sys.argv) >= 4 and sys.argv[1] == "parse":
        parse(sys.argv[2], sys.argv[3])
    else:
        sys.exit(__doc__)
