"""Host-side mirror of gkquad::single (reference: gkquad/src/single/).

Same names and call shapes as the Rust crate -- `integral(f, range)`, `Integrator::new(f)
.tolerance(..).max_evals(..).points(..).algorithm(..).run(range)`, the `Algorithm`
implementors QAGS / QAGP / AUTO, `Range`, `Points`, `IntegrationConfig` -- with one change:
the integrand is a `DeviceIntegrand` (a registry functor + a batch of parameters) instead of a
closure, and `run` integrates the whole batch on the GPU.  A batch of one returns an
`IntegrationResult` exactly like the reference; larger batches return a `BatchResult`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .common import BatchResult, IntegrationResult, Tolerance
from .engine import ALGO_AUTO, ALGO_QAG, ALGO_QAGP, ALGO_QAGS, default_engine


class Range:
    """gkquad::single::Range (single/common.rs:18-96): NaN-free by construction; an unbounded
    side is +-inf (From<RangeBounds>, :82-96)."""

    def __init__(self, begin: float, end: float):
        if math.isnan(begin) or math.isnan(end):
            raise ValueError("cannot create Range object from Range which contains NaN value.")
        self.begin, self.end = float(begin), float(end)

    @staticmethod
    def new(begin, end):
        """Range::new -> Option<Range> (single/common.rs:26-32)."""
        if math.isnan(begin) or math.isnan(end):
            return None
        return Range(begin, end)

    @staticmethod
    def from_(r) -> "Range":
        if isinstance(r, Range):
            return r
        if isinstance(r, slice):  # a..b  /  a..  /  ..b  /  ..
            a = -math.inf if r.start is None else r.start
            b = math.inf if r.stop is None else r.stop
            return Range(a, b)
        a, b = r
        return Range(-math.inf if a is None else a, math.inf if b is None else b)

    def __eq__(self, o):
        return isinstance(o, Range) and (self.begin, self.end) == (o.begin, o.end)

    def __repr__(self):
        return f"[{self.begin!r}, {self.end!r}]"


class RangeBatch:
    """Per-integral ranges: arrays a[n], b[n] (numpy or torch CUDA float64)."""

    def __init__(self, begin, end):
        self.begin, self.end = begin, end


Points = list  # gkquad::single::Points = SmallVec<[f64; 8]> (single/common.rs:12)


@dataclass
class IntegrationConfig:
    """single/common.rs:108-126 (defaults :116-125)."""
    tolerance: Tolerance = field(default_factory=Tolerance.default)
    max_evals: int = 2000
    points: list = field(default_factory=list)


class DeviceIntegrand:
    """Stands where the reference takes `F: Integrand` (single/common.rs:129-145): a functor of
    the device registry plus its parameters.  `params` is None, a list of scalars (one
    integral, or shared by a batch) or an array [nparams, n] (one column per integral)."""

    def __init__(self, functor: str, params=None, n: int | None = None):
        self.functor, self.params, self.n = functor, params, n

    def batch_size(self) -> int | None:
        p = self.params
        if p is not None and hasattr(p, "ndim") and p.ndim == 2:
            return int(p.shape[1])
        return self.n


class Algorithm:
    """trait Algorithm<F> (single/algorithm/mod.rs:16-23)."""
    ALGO = ALGO_AUTO

    def __init__(self, engine=None, workspace_capacity: int = 0):
        self._engine = engine
        # with_workspace(&mut WorkSpace::with_capacity(n)) (qags.rs:34-38): capacity = `limit`
        self.workspace_capacity = workspace_capacity

    @classmethod
    def with_workspace_capacity(cls, n: int, engine=None):
        return cls(engine, n)

    def integrate(self, f: DeviceIntegrand, range_, config: IntegrationConfig):
        eng = self._engine or default_engine()
        if isinstance(range_, RangeBatch):
            a, b = range_.begin, range_.end
        else:
            r = Range.from_(range_)
            a, b = r.begin, r.end
        n = f.batch_size()
        res = eng.integrate_batch(f.functor, a, b, f.params, algo=self.ALGO, tol=config.tolerance,
                                  max_evals=config.max_evals, points=config.points,
                                  workspace_capacity=self.workspace_capacity, n=n)
        if len(res) == 1 and n is None and not isinstance(range_, RangeBatch):
            return res[0]
        return res

    def __eq__(self, o):  # extra_traits! (single/algorithm/mod.rs:25-75): all instances equal
        return type(self) is type(o)

    def __repr__(self):
        return type(self).__name__


class QAGS(Algorithm):
    """single/algorithm/qags.rs:19-63"""
    ALGO = ALGO_QAGS


class QAGP(Algorithm):
    """single/algorithm/qagp.rs:20-64"""
    ALGO = ALGO_QAGP


class QAG(Algorithm):
    """single/algorithm/qag.rs:16-60 (deprecated in the reference since 0.0.3: "always worse than QAGS")."""
    ALGO = ALGO_QAG


class AUTO(Algorithm):
    """single/algorithm/auto.rs:7-35: QAGS when there are no points, else QAGP."""
    ALGO = ALGO_AUTO


class Integrator:
    """single/integrator.rs:31-106 -- builder methods consume and return self."""

    def __init__(self, integrand: DeviceIntegrand, algorithm: Algorithm | None = None):
        self.integrand = integrand
        self._algorithm = algorithm or AUTO()
        self.config = IntegrationConfig()

    new = classmethod(lambda cls, integrand: cls(integrand))

    @classmethod
    def with_algorithm(cls, integrand, algorithm):
        return cls(integrand, algorithm)

    def algorithm(self, algorithm: Algorithm) -> "Integrator":
        self._algorithm = algorithm
        return self

    def tolerance(self, t: Tolerance) -> "Integrator":
        assert not t.contains_nan(), "Tolerance must not contain NAN."  # integrator.rs:71
        self.config.tolerance = t
        return self

    def max_evals(self, max_evals: int) -> "Integrator":
        self.config.max_evals = int(max_evals)
        return self

    def points(self, pts) -> "Integrator":
        assert all(not math.isnan(x) for x in pts), f"cannot include NAN value in singular points: {pts}"
        self.config.points = list(pts)
        return self

    def get_algorithm(self):
        return self._algorithm

    def run(self, range_):
        return self._algorithm.integrate(self.integrand, range_, self.config)


def integral(f: DeviceIntegrand, range_):
    """single::integral (single/integral.rs:17-19): fresh QAGS + default config."""
    return QAGS().integrate(f, range_, IntegrationConfig())


def integral_with_config(f: DeviceIntegrand, range_, config: IntegrationConfig):
    """single/integral.rs:25-31: AUTO."""
    return AUTO().integrate(f, range_, config)


def _qk(rule, f: DeviceIntegrand, r):
    rr = Range.from_(r)
    out = default_engine().qk_batch(rule, f.functor, [rr.begin], [rr.end], f.params).cpu().numpy()
    return QKResult(*[float(v) for v in out[:, 0]])


@dataclass
class QKResult:
    """single/qk/mod.rs:16-25"""
    estimate: float
    delta: float
    absvalue: float
    asc: float


def qk17(f: DeviceIntegrand, r) -> QKResult:
    """single/qk/mod.rs:28-33"""
    return _qk(17, f, r)


def qk25(f: DeviceIntegrand, r) -> QKResult:
    """single/qk/mod.rs:36-41"""
    return _qk(25, f, r)


def qk33(f: DeviceIntegrand, r) -> QKResult:
    """single/qk/mod.rs:44-49"""
    return _qk(33, f, r)


def qk41(f: DeviceIntegrand, r) -> QKResult:
    """single/qk/mod.rs:52-57"""
    return _qk(41, f, r)


def qk49(f: DeviceIntegrand, r) -> QKResult:
    """single/qk/mod.rs:60-65"""
    return _qk(49, f, r)


def qk57(f: DeviceIntegrand, r) -> QKResult:
    """single/qk/mod.rs:68-73"""
    return _qk(57, f, r)
