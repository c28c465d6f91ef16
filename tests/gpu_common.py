"""Helpers shared by the GPU parity tests: run the same batch through the CUDA engine (via the
C ABI) and through the CPU oracle, and compare with the contract of BASELINE.json:north_star:
|delta value| <= max(abs_tol, rel_tol*|I|) and identical status; nevals compared and reported."""
import numpy as np

from helpers import ALGO
from oracle import oracle as O


def tol_to_oracle(t):
    return {0: ("abs", t.abs_tol), 1: ("rel", t.rel_tol), 2: ("abs_or_rel", t.abs_tol, t.rel_tol),
            3: ("abs_and_rel", t.abs_tol, t.rel_tol)}[t.kind]


def allowed_error(tol, ref_est):
    k = tol.kind
    a, r = tol.abs_tol, tol.rel_tol * np.abs(ref_est)
    if k == 0:
        return np.full_like(ref_est, a)
    if k == 1:
        return r
    return np.maximum(a, r)


def compare(gpu, ref, tol, *, exact=False, label=""):
    """gpu: BatchResult (numpy); ref: dict from the oracle.  Returns a summary dict."""
    g = gpu.numpy()
    st_mis = int((g.status != ref["status"]).sum())
    nev_mis = int((g.nevals != ref["nevals"]).sum())
    both_nan = np.isnan(g.estimate) & np.isnan(ref["estimate"])
    diff = np.where(both_nan, 0.0, np.abs(g.estimate - ref["estimate"]))
    allow = allowed_error(tol, np.where(both_nan, 0.0, ref["estimate"]))
    # an integral whose oracle status is an error carries no accuracy promise beyond delta
    ok_rows = ref["status"] == 0
    out_of_tol = int((diff[ok_rows] > allow[ok_rows] + 1e-300).sum())
    bit_equal = int(((g.estimate == ref["estimate"]) | both_nan).sum())
    dbit = int(((g.delta == ref["delta"]) | (np.isnan(g.delta) & np.isnan(ref["delta"]))).sum())
    # Status flips of transcendental integrands: the engine's exp / sin / cos differ from glibc by
    # <= 2 ulp, and QUADPACK's error estimate is a difference of two nearly equal sums, so for an
    # integral that is converged to rounding level `delta` is rounding noise and the tests
    # `delta <= tol` / `delta <= 100 eps * absvalue` can go either way (the oracle itself flips at the
    # same rate when its libm is perturbed by one ulp: tools/libm_sensitivity.py).  Such a flip is
    # "noise-level" when both sides still agree on the value within the requested tolerance.
    flips = np.nonzero(g.status != ref["status"])[0]
    noise_flips = int(sum(1 for i in flips if diff[i] <= allow[i] + 1e-300))
    s = dict(n=len(g.estimate), status_mismatch=st_mis, noise_level_flips=noise_flips,
             nevals_mismatch=nev_mis, out_of_tol=out_of_tol,
             bit_equal_estimates=bit_equal, bit_equal_deltas=dbit,
             max_ratio=float(np.max(diff[ok_rows] / np.maximum(allow[ok_rows], 1e-300))) if ok_rows.any() else 0.0,
             label=label)
    if exact:
        assert bit_equal == s["n"] and dbit == s["n"] and st_mis == 0 and nev_mis == 0, s
    return s
