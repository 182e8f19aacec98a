"""Host-side mirror of the reference's constraint functions
(/root/reference/src/TDM_Constraints.jl).  Each keeps the reference's scalar signature
(`x -> Bool` / `x -> Real`); each also carries a `.batch(X)` that evaluates a whole poll set in
the fused constraint pass of the coverage kernel (cov_eval_batch's `feasible` output).
The reference reads N, FOV and r_max from Main-scope globals; here they are arguments.
"""
from __future__ import annotations

import numpy as np

from ._lib import OPT_PROGRESSIVE_INDEX
from .engine import CoverageEngine, TAN_HALF_FOV_DEFAULT


class _Constraint:
    """A callable constraint backed by an engine configured for exactly this constraint."""

    def __init__(self, name, N, configure, device=0, progressive=False):
        self.__name__ = name
        self.N = N
        self._configure = configure
        self._device = device
        self._engine = None
        self._progressive = progressive

    # one engine per device serves every stand-alone constraint (creating a handle costs milliseconds; a
    # constraint only has to re-upload its parameters when another constraint used the engine in between)
    _shared = {}

    def engine(self) -> CoverageEngine:
        slot = _Constraint._shared.get(self._device)
        if slot is None:
            eng = CoverageEngine(self._device)
            eng.set_grid_full(1, 1, 1.0, 1.0)  # constraints do not look at the cell store
            slot = _Constraint._shared[self._device] = [eng, None]
        if slot[1] is not self:
            self._configure(slot[0])
            slot[1] = self
        return slot[0]

    def batch(self, X) -> np.ndarray:
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3 * self.N)
        if self._progressive:
            return self.engine().eval_batch(X, want_count=False, want_feasible=False,
                                            want_progressive=True)["progressive"]
        return self.engine().eval_batch(X, want_count=False)["feasible"].astype(bool)

    def __call__(self, x):
        v = self.batch(np.asarray(x, dtype=np.float64).reshape(1, -1))[0]
        return float(v) if self._progressive else bool(v)

    # what a fused objective needs to fold this constraint into its own launch
    fuse = None


def cons1(x) -> bool:
    """src/TDM_Constraints.jl:9-19 -- always true (its body is commented out upstream)."""
    return True


cons1.batch = lambda X: np.ones(np.asarray(X).reshape(len(X), -1).shape[0], dtype=bool)
cons1.fuse = {}


def create_cons3(pre_optimized_circles_MADS, FOV, d_lim, device: int = 0):
    """src/TDM_Constraints.jl:54-75 -- displacement limit per UAV, z = R / tan(FOV/2), strict `>`
    rejects.  `pre_optimized_circles_MADS`: list of Circle or [x;y;R]."""
    import math
    pre = pre_optimized_circles_MADS
    if len(pre) and hasattr(pre[0], "x"):
        pre = np.array([c.x for c in pre] + [c.y for c in pre] + [c.R for c in pre], dtype=np.float64)
    pre = np.ascontiguousarray(pre, dtype=np.float64).ravel()
    N = pre.size // 3
    t = math.tan(FOV / 2)
    d = np.ascontiguousarray(np.broadcast_to(np.asarray(d_lim, dtype=np.float64), (N,)))
    c = _Constraint("cons3", N, lambda e: e.set_params(N, np.zeros(N), 0.0, prev_xyR=pre, d_lim=d, tan_half_fov=t),
                    device)
    c.fuse = {"prev_xyR": pre, "d_lim": d, "tan_half_fov": t}
    return c


def create_cons7(N: int, FOV, device: int = 0):
    """src/TDM_Constraints.jl:142-154 -- if y < 200 then R <= 19*tan(FOV/2)."""
    import math
    t = math.tan(FOV / 2)
    c = _Constraint("cons7", N, lambda e: e.set_params(N, np.zeros(N), 0.0, tan_half_fov=t, use_cons7=True), device)
    c.fuse = {"use_cons7": True, "tan_half_fov": t}
    return c


def create_cons8(N: int, sep: float = 15.0, device: int = 0):
    """src/TDM_Constraints.jl:157-172 -- every ordered pair at horizontal distance >= 15.0."""
    c = _Constraint("cons8", N, lambda e: e.set_params(N, np.zeros(N), 0.0, sep_min=sep), device)
    c.fuse = {"sep_min": sep}
    return c


def _create_progressive(name: str, which: int, N: int, r_max, device: int):
    r = np.ascontiguousarray(r_max, dtype=np.float64).ravel()
    if which > N:
        raise ValueError(f"{name} reads UAV {which} but the swarm has {N}")

    def configure(e):
        e.set_params(N, r, 0.0)
        e.set_option(OPT_PROGRESSIVE_INDEX, which)
    return _Constraint(name, N, configure, device, progressive=True)


def create_cons1_progressive(N: int, r_max, device: int = 0):
    """src/TDM_Constraints.jl:182-195 -- sum_i max(R_i - r_max_i, 0.0)."""
    return _create_progressive("cons1_progressive", 0, N, r_max, device)


def create_cons2_progressive(N: int, r_max, device: int = 0):
    """src/TDM_Constraints.jl:197-208 -- max(R_2 - r_max_2, 0.0): the term of UAV i = 2 alone."""
    return _create_progressive("cons2_progressive", 2, N, r_max, device)


def create_cons3_progressive(N: int, r_max, device: int = 0):
    """src/TDM_Constraints.jl:210-221 -- max(R_3 - r_max_3, 0.0): the term of UAV i = 3 alone."""
    return _create_progressive("cons3_progressive", 3, N, r_max, device)
