"""BASELINE configs 1 and 5: the receding-horizon MADS loop of FullSimulation.jl on the GPU objective,
timed per MADS solve, beside the same driver on the CPU port of the reference objective (test
infrastructure: oracle/).  The reference's own recorded solve times (src/MADS_Runtime.xlsx, unknown
laptop, DirectSearch.jl): median 0.074 s."""
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import coverage_b200 as cov
from oracle import c_oracle

T = cov.TAN_HALF_FOV_DEFAULT


class CpuObjective:
    """The same closure on the CPU port (for the timing comparison only)."""

    def __init__(self, cells, N, r_max):
        self.cells, self.N, self.r_max = cells, N, r_max
        self._fused = {}

    def fuse(self, constraints):
        rest, fused = [], {}
        for c in constraints:
            f = getattr(c, "fuse", None)
            (rest.append(c) if f is None else fused.update(f))
        self._fused = fused
        return rest

    def batch(self, X, want_feasible=False):
        f = self._fused
        r = c_oracle.eval_batch(X, self.N, self.r_max, self.cells.points_of_interest.data, pre=f.get("prev_xyR"),
                                d_lim=f.get("d_lim"), tan_half_fov=f.get("tan_half_fov", T), threads=1)
        return (r["obj"], r["feasible"].astype(bool)) if want_feasible else r["obj"]

    def __call__(self, x):
        return float(self.batch(np.asarray(x)[None, :])[0])


def run(environment, steps, use_gpu):
    CF, FS = cov.CellFunctions, cov.FullSimulation
    params = FS.SimulationParameters(environment_type=environment, N_iter=100, seed=3)
    fire_rows = cov.fire_io.load_fire_rows_npz("tests/golden/fire_rows.npz") if environment == "dynamic" else None
    cells = CF.initialise_POI(CF.Cells(), environment, fire_rows=fire_rows)
    N = params.N
    cy = 330.0 if environment == "dynamic" else 250.0
    start = cov.Base_Functions.allocate_even_circles(15.0, N, 10 * T, 250.0, cy)
    r_max = params.h_max * T * np.ones(N)
    if not use_gpu:
        orig = cov.TDM_STATIC_opt.createObjective
        cov.FullSimulation.TDM_STATIC_opt.createObjective = lambda c, n, r: CpuObjective(c, n, r)
    t0 = time.perf_counter()
    try:
        inp, outp, runtimes, objs = FS.run_simulation(cells, start, cov.TDM_Constraints.cons1, [], N, r_max, params,
                                                      Nt_sim=steps)
    finally:
        if not use_gpu:
            cov.FullSimulation.TDM_STATIC_opt.createObjective = orig
    wall = time.perf_counter() - t0
    cells.close()
    return runtimes, objs, wall


if __name__ == "__main__":
    for env, steps in (("static", 40), ("dynamic", 20)):
        for use_gpu in (True, False):
            run(env, 2, use_gpu)  # warm-up
            rt, objs, wall = run(env, steps, use_gpu)
            print(f"{env:8s} {steps} steps, {'GPU objective' if use_gpu else 'CPU port     '}: median MADS solve "
                  f"{statistics.median(rt) * 1e3:7.2f} ms (min {min(rt) * 1e3:.2f}, max {max(rt) * 1e3:.2f}), "
                  f"loop wall {wall:.2f} s, last objective {objs[-1]:.1f}")
