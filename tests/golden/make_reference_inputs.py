"""Write tests/golden/reference_inputs.txt: the inputs on which tests/golden/reference_dump.jl runs the UNMODIFIED
reference (Julia) so that its outputs can pin the oracle and the CUDA path (tests/test_reference_outputs.py).

Every Float64 travels as the 16 hex digits of its bit pattern (no decimal parsing on either side).  Cases:
  static   createPOI(5.0, 5.0, 100.0, 100.0) (built by the reference itself), N = 5, 1000 candidates
  fire     rows 1..10 of src/FirePoints.xlsx (455 list entries, 254 unique cells: duplicates), N = 5, 1000 candidates
Candidates mix the MADS integer mesh (granularity 1.0) around the KAT start point, random reals, lattice ties
(centres on cell centres / cell corners with integer and half-integer radii) and near-boundary adversarial discs.

Deterministic (NumPy PCG64, fixed seeds); run from the repo root:  python tests/golden/make_reference_inputs.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
T = math.tan((100 / 180 * math.pi) / 2)


def hx(a):
    return " ".join(f"{int(v):016x}" for v in np.asarray(a, dtype=np.float64).ravel().view(np.uint64))


def candidates(rng, N, centre, spread, B, pts):
    rows = []
    x0 = np.concatenate([centre[0] + 15 * np.cos(2 * np.pi * np.arange(N) / N), centre[1] + 15 * np.sin(2 * np.pi * np.arange(N) / N),
                         np.full(N, 10 * T)])
    for b in range(B):
        kind = b % 5
        if kind == 0:    # integer mesh around the start point (what SetGranularity(p, i, 1.0) produces)
            x = np.rint(x0 + rng.normal(0, 6, 3 * N))
            x[2 * N:] = np.abs(x[2 * N:]) + 1
        elif kind == 1:  # random reals over the domain
            x = np.concatenate([rng.random(2 * N) * 500, (5 + rng.random(N) * 25) * T])
        elif kind == 2:  # random reals around the region of interest
            x = np.concatenate([centre[0] + rng.normal(0, spread, N), centre[1] + rng.normal(0, spread, N),
                                (5 + rng.random(N) * 25) * T])
        elif kind == 3:  # lattice ties: centres on cell centres / corners, integer and half-integer radii
            x = np.concatenate([np.rint((centre[0] + rng.normal(0, spread, N)) / 2.5) * 2.5,
                                np.rint((centre[1] + rng.normal(0, spread, N)) / 2.5) * 2.5,
                                np.rint(rng.random(N) * 60 + 2) / 2])
        else:            # near-boundary: a list point at distance R from the centre up to a few ulps
            x = np.empty(3 * N)
            for i in range(N):
                p = pts[rng.integers(0, len(pts))]
                R = float((5 + rng.random() * 25) * T) if rng.random() < 0.6 else float(rng.integers(4, 40))
                th = rng.random() * 2 * math.pi if rng.random() < 0.6 else float(rng.choice([0, math.pi / 2, math.pi]))
                cx, cy = p[0] - R * math.cos(th), p[1] - R * math.sin(th)
                for _ in range(int(rng.integers(0, 4))):
                    cx = math.nextafter(cx, cx + float(rng.choice([-1.0, 1.0])))
                x[i], x[N + i], x[2 * N + i] = cx, cy, R
        rows.append(x)
    return np.array(rows)


def main():
    from oracle import coverage_oracle as npo
    import importlib.util
    spec = importlib.util.spec_from_file_location("fire_io", os.path.join(ROOT, "maximumareacoverageoptimization.jl_b200", "fire_io.py"))
    fire_io = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fire_io)
    fire = np.concatenate(fire_io.load_fire_rows_npz(os.path.join(HERE, "fire_rows.npz"))[:10])
    static = npo.createPOI(5.0, 5.0, 100.0, 100.0)
    N = 5
    out = ["# inputs of tests/golden/reference_dump.jl -- written by tests/golden/make_reference_inputs.py; every Float64 is the",
           "# 16 hex digits of its bit pattern.  case <name> N <N> B <B> P <P (0: the reference builds createPOI(5,5,100,100))>"]
    for name, pts, seed, centre, spread in (("static", static, 20261018, (250.0, 250.0), 60.0),
                                            ("fire", fire, 20261019, (float(fire[:, 0].mean()), float(fire[:, 1].mean())), 40.0)):
        rng = np.random.default_rng(seed)
        B = 1000
        X = candidates(rng, N, centre, spread, B, pts)
        r_max = np.full(N, 30.0 * T)
        r_max[1] = 17.5  # not all equal: catches an index slip in the penalty / progressive terms
        pre = X[0].copy()
        d_lim = np.array([10.0, 10.0, 12.5, 10.0, 8.0])
        out.append(f"case {name} N {N} B {B} P {0 if name == 'static' else len(pts)}")
        out.append("r_max " + hx(r_max))
        out.append("pre " + hx(pre))
        out.append("d_lim " + hx(d_lim))
        if name != "static":
            for p in pts:
                out.append("point " + hx(p))
        for x in X:
            out.append("x " + hx(x))
    with open(os.path.join(HERE, "reference_inputs.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", os.path.join(HERE, "reference_inputs.txt"), len(out), "lines")


if __name__ == "__main__":
    main()
