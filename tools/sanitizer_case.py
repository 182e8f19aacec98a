"""A small but representative workload for compute-sanitizer: every kernel family once."""
import sys
import numpy as np
sys.path.insert(0, ".")
import coverage_b200 as cov
T = cov.TAN_HALF_FOV_DEFAULT
e = cov.CoverageEngine(0)
# small-swarm kernel (chunk variants), shared + unshared discs, fire list with duplicates
rows = cov.fire_io.load_fire_rows_npz("tests/golden/fire_rows.npz")
e.set_points(np.concatenate(rows[:30]), 100, 100, 5.0, 5.0)
e.set_params(5, np.full(5, 30 * T), sep_min=15.0)
X = cov.synth.random_candidates(600, 5, seed=1)
X[:100, :10] = 250 + X[:100, :10] * 0.1
r = e.eval_batch(X)
print("small kernel:", int(r["count"].sum()))
e.remove_covered(X[0]); e.add_points(rows[31])
print("argmin:", e.argmin(X[:64]))
# CTA kernel: N = 50 on a 1024^2 grid (bands), and a tiny batch routed to it
bits, n = cov.synth.fire_grid(1024)
e.set_grid_bits(bits, 1024, 1024, 500 / 1024, 500 / 1024)
e.set_params(50, np.full(50, 30 * T), sep_min=15.0)
print("cta kernel:", int(e.eval_batch(cov.synth.random_candidates(8, 50, seed=2))["count"].sum()))
e.set_option(cov.OPT_BAND_ROWS, 100)
print("cta kernel, 100-row bands:", int(e.eval_batch(cov.synth.random_candidates(4, 50, seed=2))["count"].sum()))
# brute / exact on a small grid
e.set_grid_full(64, 64, 5.0, 5.0); e.set_params(3, np.full(3, 30 * T))
for k in (cov.KERNEL_BRUTE, cov.KERNEL_EXACT):
    e.set_option(cov.OPT_KERNEL, k)
    print("kernel", k, int(e.eval_batch(cov.synth.random_candidates(64, 3, seed=3))["count"].sum()))
# fire automaton
ff = cov.DynamicArea.ForestFire(e, seed=5)
print("fire pushes:", [ff.step() for _ in range(3)])
e.close()
print("done")
