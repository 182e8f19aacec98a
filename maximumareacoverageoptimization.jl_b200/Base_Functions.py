"""The two pieces of the reference's `Base_Functions` (/root/reference/src/Base_Functions.jl) the
coverage path needs: the `Circle` layout and `allocate_even_circles` (config-1 start positions)."""
from __future__ import annotations

import math

import numpy as np

from .AreaCoverageCalculation import Circle  # noqa: F401  (src/Base_Functions.jl:37-41)


def allocate_even_circles(r_centering_cir: float, N: int, r_uav: float, center_x: float, center_y: float):
    """src/Base_Functions.jl:44-65 -- N discs evenly on a circle; returns [x;y;R]."""
    xs, ys, rs = [], [], []
    for i in range(1, N + 1):
        ref_angle = 2 * math.pi / N * (i - 1)
        xs.append(r_centering_cir * math.cos(ref_angle) + center_x)
        ys.append(r_centering_cir * math.sin(ref_angle) + center_y)
        rs.append(r_uav)
    return np.array(xs + ys + rs, dtype=np.float64)
