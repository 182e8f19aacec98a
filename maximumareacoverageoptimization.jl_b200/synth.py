"""Synthetic inputs of BASELINE.md's configurations: fire grids (a DynamicArea-style cellular
automaton, /root/reference/src/DynamicArea.jl:17-72, scaled to the grid size) and random candidate
sets.  Host-side NumPy only; the same candidates can be produced on the device by
cov_generate_candidates (Philox4x32-10), which `philox_candidates` reproduces bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

# This module is deliberately self-contained (NumPy only, no package-relative imports): bench.py's reference arm
# loads it by file path so that the CPU baseline never touches libcoverage_cuda.
TAN_HALF_FOV_DEFAULT = math.tan((100 / 180 * math.pi) / 2)  # FOV = 100/180*pi, src/FullSimulation.jl:735 (as engine.py)

GRID_SEED = 20261018
DOMAIN = 500.0


def pack_bits(fire: np.ndarray) -> np.ndarray:
    """fire[i-1, j-1] (bool, nx x ny) -> (ny, ceil(nx/32)) uint32, bit b of word w of row j-1 = cell
    i = 32w + b + 1 (include/coverage_cuda.h)."""
    nx, ny = fire.shape
    wpr = (nx + 31) // 32
    padded = np.zeros((ny, wpr * 32), dtype=np.uint8)
    padded[:, :nx] = fire.T
    b = np.packbits(padded.reshape(ny, wpr, 4, 8), axis=-1, bitorder="little").reshape(ny, wpr, 4)
    return (b[..., 0].astype(np.uint32) | (b[..., 1].astype(np.uint32) << 8) |
            (b[..., 2].astype(np.uint32) << 16) | (b[..., 3].astype(np.uint32) << 24))


def unpack_bits(bits: np.ndarray, nx: int) -> np.ndarray:
    ny, wpr = bits.shape
    by = bits.view(np.uint8).reshape(ny, wpr * 4)
    return np.unpackbits(by, axis=-1, bitorder="little")[:, :nx].T.astype(bool)


def _fire_ca_native(n: int, rng: np.random.Generator, target_frac: float, max_steps: int) -> np.ndarray:
    """EMPTY/TREE/FIRE automaton on an n x n grid: tree density 0.7 (DynamicArea.jl:20), spread
    0.5 (:21), wind 4 @ 270 deg (:47-48), ignition band at the reference's relative position
    (:11-14,35: x 200..300, y 345..355 of 500).  One draw per (cell, burning neighbour), as in
    update_grid (:52-72).  Returns the boolean burning mask [i-1, j-1]."""
    EMPTY, TREE, FIRE = 0, 1, 2
    grid = np.where(rng.random((n, n)) < 0.7, TREE, EMPTY).astype(np.uint8)
    i0, i1 = int(round(0.4 * n)), int(round(0.6 * n))
    j0, j1 = int(round(0.69 * n)), max(int(round(0.71 * n)), int(round(0.69 * n)) + 1)
    grid[i0 - 1:i1, j0 - 1:j1] = FIRE
    wind_speed, wind_direction, prob_spread = 4.0, math.radians(270.0), 0.5
    # neighbour at window index (a, b), a, b in 1..3, sits at offset (a-2, b-2)
    dirs = []
    for b in (1, 2, 3):
        for a in (1, 2, 3):
            if a == 2 and b == 2:
                continue
            p = wind_speed * math.cos(wind_direction - math.atan2(2 - b, 2 - a)) * prob_spread
            if p > 0:
                dirs.append((a - 2, b - 2, p))
    target = target_frac * n * n
    for _ in range(max_steps):
        fire = grid == FIRE
        if fire.sum() >= target:
            break
        tree = grid == TREE
        tree[0, :] = tree[-1, :] = False   # `for i in 2:n-1, j in 2:n-1`
        tree[:, 0] = tree[:, -1] = False
        new = np.zeros_like(fire)
        for da, db, p in dirs:
            nb = np.zeros_like(fire)
            # nb[i, j] = fire[i + da, j + db]
            src_i = slice(max(da, 0), n + min(da, 0))
            dst_i = slice(max(-da, 0), n + min(-da, 0))
            src_j = slice(max(db, 0), n + min(db, 0))
            dst_j = slice(max(-db, 0), n + min(-db, 0))
            nb[dst_i, dst_j] = fire[src_i, src_j]
            cand = tree & nb
            if p >= 1.0:
                new |= cand
            else:
                new |= cand & (rng.random((n, n)) < p)
        if not new.any():
            break
        grid[new] = FIRE
    return grid == FIRE


def fire_grid(n: int, seed: int = GRID_SEED, target_frac: float = 0.35, dense: bool = False):
    """Synthetic fire grid of BASELINE.md: n x n cells over the 500 m domain (dx = dy = 500/n, exact
    in binary64 for n a power of two).  n <= 512: the automaton at native resolution; larger n: the
    512 automaton upsampled, each fine cell burning iff its coarse cell burns and it holds a tree
    (density 0.7 drawn at fine resolution), rescaled so that about target_frac of all cells burn.
    Returns (bits, n_set)."""
    if dense:
        fire = np.ones((n, n), dtype=bool)
    else:
        rng = np.random.default_rng(np.random.PCG64(seed))
        base = min(n, 512)
        coarse_target = target_frac if base == n else min(0.95, target_frac / 0.7)
        coarse = _fire_ca_native(base, rng, coarse_target, max_steps=8 * base)
        if base == n:
            fire = coarse
        else:
            k = n // base
            fire = np.repeat(np.repeat(coarse, k, axis=0), k, axis=1)
            fire &= rng.random((n, n)) < 0.7
    bits = pack_bits(fire)
    return bits, int(fire.sum())


def random_candidates(B: int, N: int, seed: int, h_min: float = 5.0, h_max: float = 30.0,
                      tan_half_fov: float = TAN_HALF_FOV_DEFAULT, domain: float = DOMAIN, out=None) -> np.ndarray:
    """BASELINE.md candidate distribution: x, y ~ U(0, 500), h ~ U(5, 30), R = h*tan(50 deg);
    rows [x;y;R]; NumPy Generator(PCG64(seed))."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    X = out if out is not None else np.empty((B, 3 * N), dtype=np.float64)
    X[:, :2 * N] = rng.random((B, 2 * N)) * domain
    X[:, 2 * N:] = (h_min + rng.random((B, N)) * (h_max - h_min)) * tan_half_fov
    return X


def mesh_candidates(B: int, N: int, seed: int, granularity: float = 1.0, dtype=np.int16, h_min: float = 5.0,
                    h_max: float = 30.0, tan_half_fov: float = TAN_HALF_FOV_DEFAULT, domain: float = DOMAIN,
                    out=None) -> np.ndarray:
    """The same distribution on the MADS mesh (granularity 1.0 on every variable in the reference,
    src/TDM_STATIC_opt.jl:131-137): mesh INDICES q, uniform over the index ranges of x, y in [0, 500] and
    R in [h_min, h_max]*tan(50 deg); the candidate is q * granularity.  dtype int16 / int32 (or float32: q itself)."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    Q = out if out is not None else np.empty((B, 3 * N), dtype=dtype)
    hi = int(math.floor(domain / granularity))
    r_lo, r_hi = int(math.ceil(h_min * tan_half_fov / granularity)), int(math.floor(h_max * tan_half_fov / granularity))
    Q[:, :2 * N] = rng.integers(0, hi + 1, (B, 2 * N))
    Q[:, 2 * N:] = rng.integers(r_lo, r_hi + 1, (B, N))
    return Q


# ---- Philox4x32-10, the counter-based stream of cov_generate_candidates ----
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def _philox(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u01_53(a, b):
    v = ((a >> np.uint64(5)) << np.uint64(26)) | (b >> np.uint64(6))
    return v.astype(np.float64) * 1.1102230246251565e-16


def philox_candidates(B: int, N: int, seed: int, first_index: int = 0, lx: float = DOMAIN, ly: float = DOMAIN,
                      h_min: float = 5.0, h_max: float = 30.0, tan_half_fov: float = TAN_HALF_FOV_DEFAULT) -> np.ndarray:
    """Bit-for-bit what cov_generate_candidates writes (cov_grid_kernels.cu generate_kernel)."""
    idx = (np.arange(B, dtype=np.uint64) + np.uint64(first_index))[:, None] + np.zeros((1, N), dtype=np.uint64)
    u = np.zeros((B, 1), dtype=np.uint64) + np.arange(N, dtype=np.uint64)[None, :]
    lo, hi = idx & _MASK, idx >> np.uint64(32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    r = _philox(lo, hi, u, np.zeros_like(u), k0, k1)
    q = _philox(lo, hi, u, np.ones_like(u), k0, k1)
    X = np.empty((B, 3 * N), dtype=np.float64)
    X[:, :N] = _u01_53(r[0], r[1]) * lx
    X[:, N:2 * N] = _u01_53(r[2], r[3]) * ly
    X[:, 2 * N:] = (h_min + _u01_53(q[0], q[1]) * (h_max - h_min)) * tan_half_fov
    return X


def points_from_bits(bits: np.ndarray, nx: int, dx: float, dy: float) -> np.ndarray:
    """The reference's list layout for a bit grid: one entry per set cell, i outer / j inner like
    createPOI (src/AreaCoverageCalculation.jl:11-21)."""
    fire = unpack_bits(bits, nx)
    ii, jj = np.nonzero(fire)  # row-major over [i, j]: i outer, j inner
    pts = np.empty((ii.size, 5), dtype=np.float64)
    pts[:, 0] = (ii + 1.0) * dx - dx / 2
    pts[:, 1] = (jj + 1.0) * dy - dy / 2
    pts[:, 2] = dx * dy
    pts[:, 3] = dx * dy
    pts[:, 4] = 0.0
    return pts
