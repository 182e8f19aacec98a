#!/usr/bin/env python
"""bench.py -- coverage-objective throughput of libcoverage_cuda on BASELINE.json's workload.

Workload (config.workload): BASELINE.json configs[1] -- 5 UAVs x 1 M random candidates per step on a
256 x 256 synthetic fire grid, per GPU (weak scaling: every rank evaluates its own 1 M candidates).
One "step" = one pass of the hot path (cov_eval_batch*) over one batch of 1 M candidates.

  value  evals/s with the candidates already resident in HBM (cov_eval_batch_device), CUDA events
         on the launching stream, max over ranks.
  e2e    evals/s through the reference-facing call cov_eval_batch with pinned HOST buffers: the
         host->device copy of the candidates and the device->host copy of objective / count /
         feasibility are inside the timed region.
  roofline      HBM roofline of the coverage kernel from the ALGORITHMIC bytes per eval
                (24 N in + 8 obj + 8 count + 1 flag) and the live event-timed launch duration,
                plus the instruction-issue view the kernel is really bound by.
  cpu_baseline  oracle/coverage_oracle.c (a literal C port of the reference's Julia arithmetic;
                Julia itself is not installed) on the host cores, bounded sample, rank 0 / N=1.

`--impl reference` times that CPU port alone, all host threads, on the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UAV = 5
GRID_N = 256
B_PER_GPU = 1_000_000
SEP_MIN = 0.0
# other BASELINE configurations (parity-test cases, not the bench line): selectable for profiling only
WORKLOADS = {
    "c2": dict(n=5, grid=256, batch=1_000_000, sep=0.0,
               name="C2: 5 UAVs x 1M random candidates/step/GPU, 256x256 synthetic fire grid (BASELINE.json configs[1])"),
    "c3": dict(n=50, grid=1024, batch=65_536, sep=15.0,
               name="C3 (reduced batch): 50 UAVs x 64K candidates/step/GPU, 1024x1024 fire grid, cons8 separation"),
    "c4": dict(n=200, grid=4096, batch=8_192, sep=15.0,
               name="C4 (reduced batch): 200 UAVs x 8K candidates/step/GPU, 4096x4096 fire grid, cons8 separation"),
    "c1": dict(n=5, grid=100, batch=1_000_000, sep=0.0,
               name="C1 grid (100x100, dx=5, dense createPOI) with 1M random candidates/step/GPU"),
}
N_SETS = 4  # device-resident candidate sets rotated between steps: 4 x 120 MB > 126 MB L2
METRIC = "coverage_objective_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = "C2: 5 UAVs x 1M random candidates/step/GPU, 256x256 synthetic fire grid (BASELINE.json configs[1])"


def load_ncu_capture():
    """Numbers of the committed ncu capture of the bench kernel (profiles/), or {} for other workloads."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if N_UAV == 5 and GRID_N == 256 and B_PER_GPU == 1_000_000 and os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def load_traffic():
    """DRAM bytes per launch of the bench kernel from the committed ncu capture, or None."""
    d = load_ncu_capture()
    return d["dram_bytes_read"] + d["dram_bytes_write"] if d else None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            inside = t0 - 0.05 <= ts <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[1]))
                    power.append(float(f[3]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: use every sample we have
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local: int):
    """Pin this rank to the CPUs next to its GPU before any pinned host memory is allocated (first touch
    then places the staging pages on the GPU's NUMA node). Best effort: returns a note for the JSON line."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(local)
        bus = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return f"bound to {len(use)} CPUs local to {bus}"
        return f"no narrower local CPU set for {bus} ({len(allowed)} CPUs allowed)"
    except Exception as e:  # noqa: BLE001  (sysfs layout, permissions, torch version)
        return f"not bound ({type(e).__name__})"


def make_workload(cov):
    d = 500.0 / GRID_N
    bits, n_fire = cov.synth.fire_grid(GRID_N, dense=(GRID_N == 100))
    r_max = np.full(N_UAV, 30.0 * cov.TAN_HALF_FOV_DEFAULT)
    return bits, n_fire, d, r_max


def cpu_port_rate(cov, bits, d, r_max, seconds: float, threads: int = 0, seed: int = 12345):
    """evals/s of the CPU port (oracle/coverage_oracle.c) on a bounded sample of the workload."""
    from oracle import c_oracle
    pts = cov.synth.points_from_bits(bits, GRID_N, d, d)
    nthr = c_oracle.num_threads() if threads <= 0 else threads
    X = cov.synth.random_candidates(max(2048, 128 * nthr), N_UAV, seed=seed)
    c_oracle.eval_batch(X[:256], N_UAV, r_max, pts, sep_min=SEP_MIN, threads=threads)  # thread start-up, page-in
    t = time.perf_counter()
    c_oracle.eval_batch(X, N_UAV, r_max, pts, sep_min=SEP_MIN, threads=threads)
    rate = len(X) / (time.perf_counter() - t)
    n = int(max(len(X), min(rate * seconds, 4_000_000)))
    X = cov.synth.random_candidates(n, N_UAV, seed=seed + 1)
    t = time.perf_counter()
    out = c_oracle.eval_batch(X, N_UAV, r_max, pts, sep_min=SEP_MIN, threads=threads)
    dt = time.perf_counter() - t
    return n / dt, nthr, n, dt, out, X, pts


def run_reference(args):
    """The reference arm: the CPU port of the reference's objective on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import coverage_b200 as cov
    from oracle import c_oracle
    bits, n_fire, d, r_max = make_workload(cov)
    pts = cov.synth.points_from_bits(bits, GRID_N, d, d)
    nthr = c_oracle.num_threads()
    # one step = a bounded sample of the workload: sized for ~2 s of CPU work per step
    probe = cov.synth.random_candidates(max(256, 64 * nthr), N_UAV, seed=99)
    c_oracle.eval_batch(probe, N_UAV, r_max, pts, sep_min=SEP_MIN, threads=0)  # thread start-up, page-in
    t = time.perf_counter()
    c_oracle.eval_batch(probe, N_UAV, r_max, pts, sep_min=SEP_MIN, threads=0)
    rate = len(probe) / (time.perf_counter() - t)
    # bounded: the whole --steps/--warmup run should end within ~2 minutes
    budget_s = 100.0 / max(args.steps + args.warmup, 1)
    per_step = int(max(256, min(rate * min(2.0, budget_s), B_PER_GPU)))
    X = cov.synth.random_candidates(per_step, N_UAV, seed=1)
    for _ in range(args.warmup):
        c_oracle.eval_batch(X, N_UAV, r_max, pts, sep_min=SEP_MIN, threads=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.eval_batch(X, N_UAV, r_max, pts, sep_min=SEP_MIN, threads=0)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} of the workload's {B_PER_GPU} candidates per step, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "uavs": N_UAV, "grid": f"{GRID_N}x{GRID_N}", "fire_entries": n_fire,
                   "candidates_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample,
                         "note": "C port of the reference's Julia objective (Julia is not installed); "
                                 "pthreads over candidates"},
        "tests_per_sec": value * n_fire * N_UAV,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "span", "brute", "exact"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=24.0)
    args = ap.parse_args()
    global N_UAV, GRID_N, B_PER_GPU, WORKLOAD, SEP_MIN
    wl = WORKLOADS[args.workload]
    N_UAV, GRID_N, B_PER_GPU, WORKLOAD, SEP_MIN = wl["n"], wl["grid"], wl["batch"], wl["name"], wl["sep"]
    if args.batch > 0:
        B_PER_GPU = args.batch
    args.batch = B_PER_GPU
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import coverage_b200 as cov

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcoverage_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    N = N_UAV
    bits, n_fire, d, r_max = make_workload(cov)

    eng = cov.CoverageEngine(local)
    stream = torch.cuda.Stream(device=local)
    eng.set_stream(stream.cuda_stream)  # torch's events see the kernels on this stream
    eng.set_grid_bits(bits, GRID_N, GRID_N, d, d)
    eng.set_params(N, r_max, sep_min=SEP_MIN)
    kid = {"auto": cov.KERNEL_AUTO, "span": cov.KERNEL_SPAN, "brute": cov.KERNEL_BRUTE, "exact": cov.KERNEL_EXACT}[args.kernel]
    eng.set_option(cov.OPT_KERNEL, kid)

    row_bytes = 3 * N * 8
    # ---- device-resident candidate sets (Philox, distinct per rank and per set) ----
    dX = [eng.device_alloc(B * row_bytes) for _ in range(N_SETS)]
    d_obj, d_cnt, d_fe = eng.device_alloc(B * 8), eng.device_alloc(B * 8), eng.device_alloc(B)
    for k in range(N_SETS):
        eng.generate_candidates(dX[k], B, N, seed=1 + rank, first_index=k * B)
    eng.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(k):
        eng.eval_batch_device(dX[k % N_SETS], B, d_obj, d_cnt, d_fe)

    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            step_device(k)
        barrier()
        ms0, l0 = eng.kernel_time_total()
        launches0 = eng.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        e0.record(stream)
        for k in range(args.steps):
            step_device(k)
        e1.record(stream)
        barrier()
        t_wall1 = time.perf_counter()
        dev_ms = e0.elapsed_time(e1)
        ms1, l1 = eng.kernel_time_total()
        launches = eng.launch_count() - launches0
    kernel_ms = (ms1 - ms0) / max(l1 - l0, 1)  # average coverage-kernel launch, events inside the library
    # sanity: the device result of the last step equals a host-side recomputation on a small sample
    obj = np.empty(B)
    cnt = np.empty(B, dtype=np.int64)
    eng.memcpy_d2h(obj, d_obj)
    eng.memcpy_d2h(cnt, d_cnt)
    eng.sync()

    # ---- end to end through the host API with pinned host buffers ----
    Xh = eng.pinned((B, 3 * N))
    cov.synth.random_candidates(B, N, seed=1000 + rank, out=Xh)
    out = {"obj": eng.pinned((B,)), "count": eng.pinned((B,), np.int64), "feasible": eng.pinned((B,), np.uint8)}
    for _ in range(args.warmup):
        eng.eval_batch(Xh, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.eval_batch(Xh, out=out)  # synchronous on return: results are in host memory
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    times = torch.tensor([dev_ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, kernel_ms = (float(v) for v in times.tolist())

    if rank == 0:
        hbm_peak, peak_kind, sm_max = load_peaks()
        total = world * B * args.steps
        value = total / (dev_ms * 1e-3)
        e2e = total / (e2e_ms * 1e-3)
        bytes_per_eval = 24 * N + 8 + 8 + 1
        achieved = B * bytes_per_eval / (kernel_ms * 1e-3) / 1e9
        tests_per_eval = n_fire * N
        sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
        issue_peak = 148 * 4 * 32 * sm_mhz * 1e6  # lane-instructions/s at the clock seen under load
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "uavs": N, "grid": f"{GRID_N}x{GRID_N}", "fire_entries": n_fire,
                       "candidates_per_step_per_gpu": B, "kernel": args.kernel,
                       "l2": f"device inputs rotate over {N_SETS} x {B * row_bytes / 1e6:.0f} MB candidate sets (> 126 MB L2)",
                       "penalties": "altitude penalty 1e5*sum|R - r_max|" + (f" + cons8 separation {SEP_MIN}" if SEP_MIN > 0 else "")},
            "tests_per_sec": value * tests_per_eval,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": load_traffic(), "peak_source": f"of {peak_kind}",
                         "algorithmic_bytes_per_launch": B * bytes_per_eval, "bytes_per_eval": bytes_per_eval, "kernel": "span_small_kernel" if N <= 8 else "span_kernel", "kernel_ms": kernel_ms,
                         "note": "the kernel is instruction-issue bound, not HBM bound (DESIGN.md); see issue"},
            "issue": {"algorithmic_tests_per_sec_per_gpu": B * tests_per_eval / (kernel_ms * 1e-3),
                      "lane_instr_peak_per_sec": issue_peak,
                      "brute_force_ceiling_tests_per_sec": issue_peak / 6.0,
                      "frac_of_brute_force_ceiling": B * tests_per_eval / (kernel_ms * 1e-3) / (issue_peak / 6.0),
                      "ncu_issue_slot_utilisation": load_ncu_capture().get("issue_slot_utilisation"),
                      "ncu_warp_instructions_per_candidate": load_ncu_capture().get("warp_instructions_per_candidate")},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * row_bytes, "d2h_bytes_per_step": B * 17,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "host_placement": numa_note,
            "clocks": clocks,
            "check": {"count_sum_last_step": int(cnt.sum()), "obj_finite": bool(np.isfinite(obj).all())},
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, nthr, n, dt, ref, Xc, pts = cpu_port_rate(cov, bits, d, r_max, args.cpu_seconds)
            got = eng.eval_batch(Xc)
            rate1, _, n1, dt1, _, _, _ = cpu_port_rate(cov, bits, d, r_max, min(3.0, args.cpu_seconds / 4), threads=1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": nthr, "kind": "port",
                                    "sample": f"{n} candidates of the same workload in {dt:.1f} s",
                                    "value_1_thread": rate1, "sample_1_thread": f"{n1} candidates in {dt1:.1f} s",
                                    "parity_on_sample": bool(np.array_equal(got["count"], ref["count"]) and
                                                             np.array_equal(got["obj"], ref["obj"]))}
        print(json.dumps(line))
    for p in dX + [d_obj, d_cnt, d_fe]:
        eng.device_free(p)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
