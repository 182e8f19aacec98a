"""Host-side mirror of the reference's forest-fire generator (/root/reference/src/DynamicArea.jl),
running the automaton on the device and feeding the coverage cell store directly -- no xlsx
hand-off.

The reference script draws from Julia's global RNG; here every draw is a counter-based Philox
uniform keyed by (seed, step, cell, neighbour), so host and device restatements agree bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

from .engine import CoverageEngine

EMPTY, TREE, FIRE = 0, 1, 2  # src/DynamicArea.jl:17


class ForestFire:
    """Parameters as in src/DynamicArea.jl:5-22,47-48; `engine` holds the CA state and the cell store."""

    def __init__(self, engine: CoverageEngine, dx=5.0, dy=5.0, X=500.0, Y=500.0, x_start1=200.0, x_start2=300.0,
                 y_start1=345.0, y_start2=355.0, forest_density=0.7, prob_spread=0.5, wind_speed=4.0,
                 wind_direction=math.radians(270.0), seed=0, push_initial=True):
        self.engine = engine
        self.dx, self.dy = float(dx), float(dy)
        self.nx, self.ny = int(round(X / dx)), int(round(Y / dy))
        self.prob_spread, self.wind_speed, self.wind_direction = prob_spread, wind_speed, wind_direction
        self.seed = int(seed)
        self.t = 0
        rng = np.random.default_rng(np.random.PCG64(seed))
        # grid[i, j]: TREE with probability forest_density, else EMPTY (:26-33); i outer, j inner
        grid = np.where(rng.random((self.nx, self.ny)) < forest_density, TREE, EMPTY).astype(np.uint8)
        i0, i1 = int(round(x_start1 / dx)), int(round(x_start2 / dx))
        j0, j1 = int(round(y_start1 / dy)), int(round(y_start2 / dy))
        grid[i0 - 1:i1, j0 - 1:j1] = FIRE  # :35 (1-based inclusive ranges)
        self.initial_grid = grid
        engine.fire_init(grid.T.ravel(), self.nx, self.ny, self.dx, self.dy, push_initial=push_initial)

    def step(self, append: bool = True) -> int:
        """One update_grid() (:52-72): returns the number of list entries pushed."""
        self.t += 1
        return self.engine.fire_step(self.wind_speed, self.wind_direction, self.prob_spread, self.seed, self.t,
                                     append=append)

    def grid(self) -> np.ndarray:
        """Current state as grid[i-1, j-1]."""
        return self.engine.fire_state().reshape(self.ny, self.nx).T
