"""Shared helpers for the test-suite (not product code)."""
import math

ALGO = {"AUTO": 0, "QAGS": 1, "QAGP": 2, "QAG": 3}
STATUS = {"None": 0, "InsufficientIteration": 1, "RoundoffError": 2, "SubrangeTooSmall": 3,
          "Divergent": 4, "NanValueEncountered": 5}


def num(v):
    """JSON cannot hold infinities: the fixtures spell them 'inf' / '-inf'."""
    if isinstance(v, str):
        return {"inf": math.inf, "-inf": -math.inf}[v]
    return float(v)


def assert_rel(left, right, tol):
    """The reference's assert_rel! macro (tests/common/macros.rs:1-22): relative error; when the
    expected value is 0 the check is `left < tol` (no abs)."""
    if right == 0.0:
        assert left < tol, (left, right)
    else:
        assert abs(left - right) / abs(right) <= tol, (left, right)
