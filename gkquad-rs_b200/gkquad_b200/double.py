"""Host-side mirror of gkquad::double (reference: gkquad/src/double/)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field

from .common import Tolerance
from .engine import ALGO_AUTO, ALGO_QAG, ALGO_QAGP, ALGO_QAGS, default_engine
from .single import Range


class Rectangle:
    """double/range.rs:11-34.  Rectangle.new panics on infinite sides (:20-22); the tuple
    conversion (From<(R1, R2)>, :27-34) allows them."""

    def __init__(self, xrange: Range, yrange: Range):
        self.xrange, self.yrange = xrange, yrange

    @staticmethod
    def new(x1, x2, y1, y2):
        if any(math.isnan(v) for v in (x1, x2, y1, y2)):
            return None
        if not all(math.isfinite(v) for v in (x1, x2, y1, y2)):
            raise ValueError("Infinite interval in 2-dimension is not already supported.")
        return Rectangle(Range(x1, x2), Range(y1, y2))

    @staticmethod
    def from_(r):
        return Rectangle(Range.from_(r[0]), Range.from_(r[1]))


class DynamicY:
    """double/range.rs:77-112: x in [x1, x2], y in yrange(x).  The closure becomes a registry
    y-range functor: name + parameters ('const' (y0, y1) | 'zero_to_x' | 'circle' (R^2))."""

    def __init__(self, x1, x2, yrange: str, yparams=None):
        self.xrange = Range(x1, x2)
        self.yrange, self.yparams = yrange, yparams

    new = classmethod(lambda cls, x1, x2, yrange, yparams=None: cls(x1, x2, yrange, yparams))


class DynamicX:
    """double/range.rs:38-75: y in [y1, y2], x in xrange(y).  The algorithms swap the arguments of
    the integrand, integrate over y outside and x inside, and swap the coordinates of the singular
    points (double/algorithm/qags.rs:29-44, qagp.rs:34-53)."""

    def __init__(self, xrange: str, y1, y2, xparams=None):
        self.yrange = Range(y1, y2)
        self.xrange, self.xparams = xrange, xparams

    new = classmethod(lambda cls, xrange, y1, y2, xparams=None: cls(xrange, y1, y2, xparams))


def _into_range2(r):
    """IntoRange2 (double/range.rs:125-173)."""
    if isinstance(r, (DynamicY, DynamicX)):
        return r
    if isinstance(r, Rectangle):
        return DynamicY(r.xrange.begin, r.xrange.end, "const", [r.yrange.begin, r.yrange.end])  # qags.rs:18-27
    return _into_range2(Rectangle.from_(r))


@dataclass
class IntegrationConfig2:
    """double/common.rs:14-32"""
    tolerance: Tolerance = field(default_factory=Tolerance.default)
    max_evals: int = 100000
    points: list = field(default_factory=list)


class DeviceIntegrand2:
    """Stands where the reference takes `F: Integrand2` (double/common.rs:35-45)."""

    def __init__(self, functor: str, params=None, n: int | None = None):
        self.functor, self.params, self.n = functor, params, n


class Algorithm2:
    """trait Algorithm2<F, R> (double/algorithm/mod.rs:7-10)."""
    ALGO = ALGO_AUTO

    def __init__(self, engine=None):
        self._engine = engine

    def integrate(self, f: DeviceIntegrand2, range_, config: IntegrationConfig2):
        eng = self._engine or default_engine()
        r = _into_range2(range_)
        n = f.n
        p = f.params
        if p is not None and hasattr(p, "ndim") and p.ndim == 2:
            n = int(p.shape[1])
        if isinstance(r, DynamicX):
            res = eng.integrate2_batch(f.functor, r.xrange, r.yrange.begin, r.yrange.end, fparams=p,
                                       yparams=r.xparams, algo=self.ALGO, tol=config.tolerance,
                                       max_evals=config.max_evals, n=n, swap_xy=True,
                                       points=[(y, x) for (x, y) in config.points])  # qagp.rs:49
        else:
            res = eng.integrate2_batch(f.functor, r.yrange, r.xrange.begin, r.xrange.end, fparams=p,
                                       yparams=r.yparams, algo=self.ALGO, tol=config.tolerance,
                                       max_evals=config.max_evals, points=config.points, n=n)
        return res[0] if (len(res) == 1 and n is None) else res

    def __eq__(self, o):
        return type(self) is type(o)

    def __repr__(self):
        return type(self).__name__


class QAGS2(Algorithm2):
    """double/algorithm/qags.rs:9-97"""
    ALGO = ALGO_QAGS


class QAGP2(Algorithm2):
    """double/algorithm/qagp.rs:13-139"""
    ALGO = ALGO_QAGP


class QAG2(Algorithm2):
    """double/algorithm/qag.rs:9-100 (deprecated in the reference)."""
    ALGO = ALGO_QAG


class AUTO2(Algorithm2):
    """double/algorithm/auto.rs:6-36"""
    ALGO = ALGO_AUTO


class Integrator2:
    """double/integrator.rs:9-88"""

    def __init__(self, integrand: DeviceIntegrand2, algorithm: Algorithm2 | None = None):
        self.integrand = integrand
        self._algorithm = algorithm or AUTO2()
        self.config = IntegrationConfig2()

    new = classmethod(lambda cls, integrand: cls(integrand))

    @classmethod
    def with_algorithm(cls, integrand, algorithm):
        return cls(integrand, algorithm)

    def algorithm(self, algorithm):
        self._algorithm = algorithm
        return self

    def tolerance(self, t: Tolerance):
        assert not t.contains_nan(), "Tolerance must not contain NAN."
        self.config.tolerance = t
        return self

    def max_evals(self, max_evals: int):
        self.config.max_evals = int(max_evals)
        return self

    def points(self, pts):
        assert all(not math.isnan(x) and not math.isnan(y) for x, y in pts), \
            f"cannot include NAN value in singular points: {pts}"
        self.config.points = [tuple(p) for p in pts]
        return self

    def get_algorithm(self):
        return self._algorithm

    def run(self, range_):
        return self._algorithm.integrate(self.integrand, range_, self.config)


def integral2(f: DeviceIntegrand2, r):
    """double/integral.rs:9-16: QAGS2 + default config."""
    return QAGS2().integrate(f, r, IntegrationConfig2())


def integral2_with_config(f: DeviceIntegrand2, r, config: IntegrationConfig2):
    """double/integral.rs:18-31: AUTO2."""
    return AUTO2().integrate(f, r, config)
