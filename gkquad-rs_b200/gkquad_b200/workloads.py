"""Synthetic parameterised-integrand batches for the BASELINE.json configurations
(SURVEY.md section 8d).  Parameters come from a counter-based generator so that the GPU
engine, every rank, the CPU baseline and a reader can regenerate identical inputs:

    u_k(i) = (splitmix64(seed + 2 i + k) >> 11) * 2^-53,   seed = 0x9E3779B97F4A7C15
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .common import Tolerance

SEED = 0x9E3779B97F4A7C15
_M = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(k: int, start: int, count: int) -> np.ndarray:
    """u_k(i) for i in [start, start+count)."""
    i = np.arange(start, start + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        ctr = np.uint64(SEED) + np.uint64(2) * i + np.uint64(k)
    return (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


@dataclass
class Workload:
    name: str
    dim: int
    functor: str
    algo: str                      # "QAGS" | "QAGP" | "AUTO"
    tol: Tolerance
    max_evals: int
    a: float
    b: float
    params: object = None          # callable(start, count) -> [np, count] array, or None
    points: object = None          # callable(params) -> (pts_offsets, pts) | list (shared) | None
    yrange: str = "const"
    yparams: list = field(default_factory=list)
    truth: object = None           # callable(params) -> exact values, or None
    description: str = ""

    def make_params(self, start: int, count: int):
        return None if self.params is None else np.ascontiguousarray(self.params(start, count))


def _s1(lo0, w0, lo1, w1):
    return lambda s, c: np.stack([lo0 + w0 * uniform(0, s, c), lo1 + w1 * uniform(1, s, c)])


def _truth_expcos(p):
    p0, p1 = p
    return (p0 * (1 - np.exp(-p0) * np.cos(p1)) + p1 * np.exp(-p0) * np.sin(p1)) / (p0 ** 2 + p1 ** 2)


REL10 = Tolerance.Relative(1e-10)
INF = math.inf

WORKLOADS = {
    # cfg-2: 1M parameterised smooth integrands, QAGS, Relative(1e-10)  (headline = S1)
    "S1": Workload("S1", 1, "expcos_p", "QAGS", REL10, 2000, 0.0, 1.0, _s1(0.5, 1.5, 1.0, 9.0),
                   truth=_truth_expcos,
                   description="exp(-p0 x) cos(p1 x) on [0,1], p0 in [0.5,2], p1 in [1,10]"),
    "S2": Workload("S2", 1, "lorentz_p", "QAGS", REL10, 2000, 0.0, 1.0,
                   lambda s, c: (1.0 + 19.0 * uniform(0, s, c))[None, :],
                   truth=lambda p: np.arctan(np.sqrt(p[0])) / np.sqrt(p[0]),
                   description="1/(1 + p0 x^2) on [0,1], p0 in [1,20]"),
    "S3": Workload("S3", 1, "expcos_p", "QAGS", REL10, 2000, 0.0, 1.0, _s1(0.5, 2.5, 5.0, 35.0),
                   truth=_truth_expcos, description="S1 with p0 in [0.5,3], p1 in [5,40] (stress)"),
    # cfg-3: QAGP with user singular points
    "P0": Workload("P0", 1, "recip_p", "AUTO", Tolerance.Absolute(1.49e-8), 2000, -1.0, 1.0,
                   lambda s, c: (0.5 + 1.5 * uniform(0, s, c))[None, :], points=[0.0],
                   truth=lambda p: np.zeros_like(p[0]),
                   description="p0/x on [-1,1], points [0] (benches/single.rs:19-29 parameterised)"),
    "P1": Workload("P1", 1, "abspow_p", "QAGP", REL10, 2000, 0.0, 1.0,
                   lambda s, c: np.stack([0.2 + 0.6 * uniform(0, s, c), 0.1 + 0.5 * uniform(1, s, c)]),
                   points="param0",
                   truth=lambda p: (p[0] ** (1 - p[1]) + (1 - p[0]) ** (1 - p[1])) / (1 - p[1]),
                   description="|x-c|^-p on [0,1], point [c]"),
    "P2": Workload("P2", 1, "logabs_p", "QAGP", REL10, 2000, 0.0, 1.0,
                   lambda s, c: (0.2 + 0.6 * uniform(0, s, c))[None, :], points="param0",
                   truth=lambda p: p[0] * np.log(p[0]) + (1 - p[0]) * np.log(1 - p[0]) - 1,
                   description="ln|x-c| on [0,1], point [c]"),
    # cfg-4: infinite range through the variable transform
    "I0": Workload("I0", 1, "lorentz_p", "AUTO", Tolerance.default(), 2000, -INF, INF,
                   lambda s, c: np.ones((1, c)), truth=lambda p: np.full_like(p[0], math.pi),
                   description="1/(1+x^2) on (-inf,inf), defaults (benches/single.rs:31-39)"),
    "I1": Workload("I1", 1, "lorentz_p", "QAGS", REL10, 2000, -INF, INF,
                   lambda s, c: (0.5 + 2.5 * uniform(0, s, c))[None, :],
                   truth=lambda p: math.pi / np.sqrt(p[0]),
                   description="1/(1 + p0 x^2) on (-inf,inf)"),
    "I2": Workload("I2", 1, "gauss_p", "QAGS", REL10, 2000, -INF, INF,
                   lambda s, c: (0.5 + 2.5 * uniform(0, s, c))[None, :],
                   truth=lambda p: np.sqrt(math.pi / p[0]), description="exp(-p0 x^2) on (-inf,inf)"),
    # cfg-5: double integral with singular points
    "D0": Workload("D0", 2, "recip_xy_p", "AUTO", Tolerance.Absolute(1.49e-8), 100000, -1.0, 1.0,
                   lambda s, c: (0.5 + 1.5 * uniform(0, s, c))[None, :], points=[(0.0, 0.0)],
                   yrange="const", yparams=[-1.0, 1.0], truth=lambda p: np.zeros_like(p[0]),
                   description="p0/(x y) on [-1,1]^2, point (0,0) (benches/double.rs:18-28)"),
    "D1": Workload("D1", 2, "d1_p", "QAGP", Tolerance.Relative(1e-6), 100000, -1.0, 1.0,
                   lambda s, c: np.stack([-0.5 + uniform(0, s, c), -0.5 + uniform(1, s, c)]),
                   points="param01", yrange="const", yparams=[-1.0, 1.0],
                   description="1/sqrt(|x-cx|+|y-cy|) on [-1,1]^2, point (cx,cy)"),
}


def csr_points_from_param0(params: np.ndarray):
    """Per-integral single break point c = params[0][i] in CSR form."""
    n = params.shape[1]
    return np.arange(n + 1, dtype=np.int64), np.ascontiguousarray(params[0])


def block_cyclic_shard(n_total: int, rank: int, world: int, block: int = 4096):
    """Indices of the integrals owned by `rank` (i -> (i / B) mod G)."""
    nb = (n_total + block - 1) // block
    mine = np.arange(rank, nb, world, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n_total]
