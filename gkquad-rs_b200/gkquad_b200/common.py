"""Host-side mirror of gkquad's crate-level types (reference: gkquad/src/common.rs, error.rs).

Names, argument meaning and error behaviour follow the Rust crate so that code (and tests)
written against gkquad read the same here; the only difference is that results come in batches.
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass

import numpy as np


class RuntimeError_(enum.IntEnum):
    """gkquad::RuntimeError (error.rs:140-178); numbering = declaration order, 0 = no error."""
    InsufficientIteration = 1
    RoundoffError = 2
    SubrangeTooSmall = 3
    Divergent = 4
    NanValueEncountered = 5


# the reference exports it as `RuntimeError`; keep Python's builtin usable too
RuntimeError = RuntimeError_


class Tolerance:
    """gkquad::Tolerance (common.rs:9-49)."""
    ABSOLUTE, RELATIVE, ABS_OR_REL, ABS_AND_REL = 0, 1, 2, 3

    def __init__(self, kind: int, abs_tol: float = 0.0, rel_tol: float = 0.0):
        self.kind, self.abs_tol, self.rel_tol = int(kind), float(abs_tol), float(rel_tol)

    @classmethod
    def Absolute(cls, x):
        return cls(cls.ABSOLUTE, x, 0.0)

    @classmethod
    def Relative(cls, y):
        return cls(cls.RELATIVE, 0.0, y)

    @classmethod
    def AbsOrRel(cls, x, y):
        return cls(cls.ABS_OR_REL, x, y)

    @classmethod
    def AbsAndRel(cls, x, y):
        return cls(cls.ABS_AND_REL, x, y)

    @classmethod
    def default(cls):
        return cls.AbsOrRel(1.49e-8, 1.49e-8)  # common.rs:44-49

    def contains_nan(self) -> bool:  # common.rs:22-31
        if self.kind == self.ABSOLUTE:
            return math.isnan(self.abs_tol)
        if self.kind == self.RELATIVE:
            return math.isnan(self.rel_tol)
        return math.isnan(self.abs_tol) or math.isnan(self.rel_tol)

    def to_abs(self, value):  # common.rs:34-41
        if self.kind == self.ABSOLUTE:
            return self.abs_tol + 0 * value
        if self.kind == self.RELATIVE:
            return value * self.rel_tol
        if self.kind == self.ABS_OR_REL:
            return np.fmax(self.abs_tol, value * self.rel_tol)
        return np.fmin(self.abs_tol, value * self.rel_tol)

    def __eq__(self, o):
        return isinstance(o, Tolerance) and (self.kind, self.abs_tol, self.rel_tol) == (o.kind, o.abs_tol, o.rel_tol)

    def __repr__(self):
        n = ["Absolute", "Relative", "AbsOrRel", "AbsAndRel"][self.kind]
        args = {0: f"{self.abs_tol}", 1: f"{self.rel_tol}"}.get(self.kind, f"{self.abs_tol}, {self.rel_tol}")
        return f"Tolerance.{n}({args})"


@dataclass
class Solution:
    """gkquad::Solution (common.rs:213-231)."""
    estimate: float = 0.0
    delta: float = 1.7976931348623157e308
    nevals: int = 0


class IntegrationResult:
    """ValueWithError<Solution, RuntimeError> (common.rs:60-209): the partial value is always
    present; `error` is None or a RuntimeError."""

    def __init__(self, value: Solution, error=None):
        self.value, self.error = value, error

    def has_err(self):
        return self.error is not None

    def err(self):
        return self.error

    def ok(self):
        return None if self.error is not None else self.value

    def unwrap(self) -> Solution:  # panics in the reference (common.rs:193-208)
        if self.error is not None:
            raise ValueError(f"called `unwrap()` on a result with error: {self.error.name}")
        return self.value

    def unwrap_unchecked(self) -> Solution:
        return self.value

    def __repr__(self):
        return f"IntegrationResult(value={self.value}, error={self.error.name if self.error else None})"


class BatchResult:
    """Structure-of-arrays results of a batch call: estimate, delta (float64), nevals, status
    (uint32); numpy arrays (host entry) or torch CUDA tensors (device entry)."""

    def __init__(self, estimate, delta, nevals, status):
        self.estimate, self.delta, self.nevals, self.status = estimate, delta, nevals, status

    def __len__(self):
        return int(self.estimate.shape[0])

    def numpy(self) -> "BatchResult":
        def cv(t):
            return t if isinstance(t, np.ndarray) else t.detach().cpu().numpy()
        nev, st = cv(self.nevals), cv(self.status)
        return BatchResult(cv(self.estimate), cv(self.delta), nev.view(np.uint32) if nev.dtype != np.uint32 else nev,
                           st.view(np.uint32) if st.dtype != np.uint32 else st)

    def __getitem__(self, i) -> IntegrationResult:
        r = self.numpy()
        st = int(r.status[i])
        return IntegrationResult(Solution(float(r.estimate[i]), float(r.delta[i]), int(r.nevals[i])),
                                 RuntimeError_(st) if st else None)

    def has_err(self):
        return bool((self.numpy().status != 0).any())

    def unwrap(self):
        r = self.numpy()
        bad = np.nonzero(r.status)[0]
        if bad.size:
            raise ValueError(f"{bad.size} integrals failed; first: index {bad[0]} "
                             f"{RuntimeError_(int(r.status[bad[0]])).name}")
        return r
