"""Property test of the span certificate (csrc/cov_span_common.cuh fast_span) on the CPU.

The kernels accept a row's span [lo, hi] from FP32 arithmetic only when a certificate holds; otherwise the exact FP64
walk decides.  The GPU tests check the outcome (bit-exact counts); this test checks the CERTIFICATE ITSELF, away from
any GPU: a NumPy float32 emulation of make_sdisc + fast_span on hundreds of thousands of random and adversarial
(disc, row) pairs, against the reference predicate evaluated in Float64 on every column of the row
(src/AreaCoverageCalculation.jl:70: sqrt(fl(fl(ddx^2) + fl(ddy^2))) < R).  Whenever the emulation says "certainly
[lo, hi]" or "certainly empty", the Float64 truth must agree; it may say "slow path" as often as it likes (and the
test reports how often, to keep the band honest).

The emulation is deliberately a little WORSE than the device: the reciprocal square root is perturbed by up to 2 ulp
(rsqrt.approx may be off by that much) and the fused multiply-adds are emulated through float64 (double rounding),
so the margins of the certificate, not luck, have to carry the result.
"""
import math

import numpy as np

f32 = np.float32


def thresholds(orc, R):
    return np.array([orc.threshold_by_search(float(r)) for r in R])


def exact_spans(cx, cy, T, j, nx, dx, dy):
    """Covered columns of row j (1-based) per case, in Float64 exactly as the reference forms the radicand."""
    i = np.arange(1, nx + 1, dtype=np.float64)[None, :]
    px = i * dx - dx / 2
    py = (j.astype(np.float64) * dy - dy / 2)[:, None]
    ddx = px - cx[:, None]
    ddy = py - cy[:, None]
    s = ddx * ddx + ddy * ddy
    inside = s < T[:, None]
    any_in = inside.any(axis=1)
    lo = np.where(any_in, inside.argmax(axis=1) + 1, 1)
    hi = np.where(any_in, nx - inside[:, ::-1].argmax(axis=1), 0)
    # the covered set of a row is one interval: check that too
    assert np.array_equal(inside.sum(axis=1), np.where(any_in, hi - lo + 1, 0))
    return lo, hi


def fma32(a, b, c):
    """float32 fma emulated through float64 (exact product, two roundings)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def emulate(cx, cy, R, T, j, nx, ny, dx, dy, rng):
    """make_sdisc + fast_span in float32.  Returns (status, lo, hi): status 0 empty, 1 span, 2 slow."""
    inv_dx, inv_dy = 1.0 / dx, 1.0 / dy
    dxf, dyf, inv_dxf = f32(dx), f32(dy), f32(inv_dx)
    extent = np.nextafter(f32(max(nx * dx, ny * dy)), f32(np.inf))
    k_ca = np.nextafter(f32(0.75 * inv_dx * 1.000001), f32(np.inf))
    gx, gy = cx * inv_dx + 0.5, cy * inv_dy + 0.5
    ic, jc = np.rint(gx), np.rint(gy)
    fx = (cx - (ic * dx - dx / 2)).astype(f32)
    fy = (cy - (jc * dy - dy / 2)).astype(f32)
    Tf = T.astype(f32)
    rt = np.sqrt(Tf)
    M = np.maximum(np.maximum(np.abs(cx.astype(f32)), np.abs(cy.astype(f32))), extent) * f32(1.0000002)
    E = f32(1.1920929e-07) * (rt + np.maximum(np.abs(fx), np.abs(fy))) + f32(3.5527137e-15) * M
    delta = f32(2.0) * (f32(4.0) * rt * E + f32(2.0) * E * E + Tf * f32(4.76837158203125e-07))
    regular = (np.abs(gx) < 4194304.0) & (np.abs(gy) < 4194304.0) & (T < 1e30) & (Tf > f32(64.0) * E * E) & (Tf > delta)
    icf, jcf = ic.astype(f32), jc.astype(f32)
    # fast_span
    v = j.astype(f32) - jcf
    y = fma32(v, np.full_like(v, dyf), -fy)
    dy2 = y * y
    w2 = Tf - dy2
    with np.errstate(invalid="ignore", divide="ignore"):
        rs = (f32(1.0) / np.sqrt(w2.astype(np.float64))).astype(f32)
        # rsqrt.approx: up to ~2 ulp off
        rs = rs * (f32(1.0) + rng.integers(-2, 3, size=rs.shape).astype(f32) * f32(1.1920929e-07))
        w = w2 * rs
        xl = (fx - w) * inv_dxf
        xr = (fx + w) * inv_dxf
        ulo, uhi = np.ceil(xl), np.floor(xr)
        e = fma32(delta * k_ca, rs, np.maximum(np.abs(xl), np.abs(xr)) * f32(1.3487e-06))
        a, b = ulo - xl, xr - uhi
        w2min = f32(10.0) * delta
        deep = w2 > w2min
        sure = deep & (np.minimum(a, b) > e) & (np.maximum(a, b) < f32(1.0) - e)
        nxf = f32(nx)
        lof = np.fmax(icf + ulo, f32(1.0))
        hif = np.fmin(icf + uhi, nxf)
    st = np.where(sure, np.where(lof <= hif, 1, 0), 2)
    st = np.where(w2 < -w2min, 0, st)
    st = np.where(regular, st, 2)
    lo = np.where(st == 1, lof, 1).astype(np.int64)
    hi = np.where(st == 1, hif, 0).astype(np.int64)
    return st, lo, hi


def run_case(orc, rng, n_cases, nx, ny, dx, dy, adversarial):
    R = np.where(rng.random(n_cases) < 0.7, (5 + rng.random(n_cases) * 25) * math.tan(math.radians(50)),
                 rng.integers(1, 60, n_cases) * 0.5).astype(np.float64) * (dx / (500.0 / 256))
    cx = rng.random(n_cases) * nx * dx * 1.2 - 0.1 * nx * dx
    cy = rng.random(n_cases) * ny * dy * 1.2 - 0.1 * ny * dy
    j = np.clip(np.rint(cy / dy + 0.5 + (rng.random(n_cases) * 2 - 1) * (R / dy + 1)), 1, ny).astype(np.int64)
    if adversarial:
        # put a lattice point of the row at distance R (up to a few ulps) from the centre: the boundary of the span
        # falls on a cell centre, where a naive FP32 decision would be a coin toss
        i = rng.integers(1, nx + 1, n_cases)
        px, py = i * dx - dx / 2, j * dy - dy / 2  # noqa: F841
        ddy = py - cy
        ok = np.abs(ddy) < R
        w = np.sqrt(np.where(ok, R * R - ddy * ddy, 0.0))
        cx = np.where(ok, px - np.where(rng.random(n_cases) < 0.5, w, -w), cx)
        for _ in range(3):
            cx = np.where(rng.random(n_cases) < 0.5, np.nextafter(cx, cx + rng.choice([-1.0, 1.0], n_cases)), cx)
        if adversarial == "near":
            # ... and then move it away by 1e-8 ... 1e-2 cells (log-uniform): some of these fall inside the zone the
            # certificate refuses, some just outside it, where it must be right
            cx = cx + rng.choice([-1.0, 1.0], n_cases) * dx * 10.0 ** (-8.0 + 6.0 * rng.random(n_cases))
    T = thresholds(orc, R)
    st, lo, hi = emulate(cx, cy, R, T, j, nx, ny, dx, dy, rng)
    elo, ehi = exact_spans(cx, cy, T, j, nx, dx, dy)
    empty_exact = ehi < elo
    span = st == 1
    assert np.array_equal(lo[span], elo[span]) and np.array_equal(hi[span], ehi[span]), "a certified span is wrong"
    assert not (span & empty_exact).any()
    assert empty_exact[st == 0].all(), "a certified empty row is not empty"
    return float((st == 2).mean()), float(span.mean())


def test_span_certificate_never_certifies_a_wrong_span(orc):
    rng = np.random.default_rng(20261018)
    slow = {}
    for name, nx, ny, dx, dy, n, adv in (
            ("C2 lattice, random", 256, 256, 500 / 256, 500 / 256, 150_000, False),
            ("C2 lattice, boundary on a cell centre", 256, 256, 500 / 256, 500 / 256, 60_000, True),
            ("C2 lattice, boundary near a cell centre", 256, 256, 500 / 256, 500 / 256, 250_000, "near"),
            ("C1 lattice (dx = 5), ties", 100, 100, 5.0, 5.0, 40_000, True),
            ("C1 lattice (dx = 5), near ties", 100, 100, 5.0, 5.0, 100_000, "near"),
            ("non-FP32-exact pitch 0.1 x 0.3", 77, 45, 0.1, 0.3, 60_000, "near"),
            ("coarse pitch 1e3", 64, 64, 1.0e3, 1.0e3, 40_000, "near"),
            ("fine pitch 1e-3", 200, 150, 1.0e-3, 1.0e-3, 40_000, "near")):
        s, sp = run_case(orc, rng, n, nx, ny, dx, dy, adv)
        slow[name] = (s, sp)
    # the band must stay narrow on the bench lattice (a wide one would be safe and slow) ...
    assert slow["C2 lattice, random"][0] < 0.002 and slow["C2 lattice, random"][1] > 0.5, slow
    # ... an exact tie must never be certified, and the near-tie sets must exercise both outcomes
    assert slow["C2 lattice, boundary on a cell centre"][1] == 0.0, slow
    for k in ("C2 lattice, boundary near a cell centre", "C1 lattice (dx = 5), near ties"):
        assert 0.05 < slow[k][0] < 0.95 and slow[k][1] > 0.05, slow
