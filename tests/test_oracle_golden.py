"""Pins the CPU oracle (oracle/) to the reference's own golden vectors.

Mirrors the reference's test harnesses: tests/qk_single.rs:14-36 (rule level, forward and
reversed range), tests/algorithms_single.rs:45-81 (driver level incl. exact nevals, status and
`order`, forward and reversed) and tests/algorithms_double.rs:19-42, at the reference's tolerances.
CPU-only: runs under `-m "not gpu"`.
"""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from helpers import ALGO, STATUS, assert_rel, num
from oracle import oracle as O

G = json.load(open(Path(__file__).parent / "golden" / "reference_vectors.json"))


@pytest.mark.parametrize("case", G["qk"], ids=lambda c: c["name"])
def test_qk_rule(case):
    r = O.qk(case["rule"], case["functor"], case["a"], case["b"], case["params"])
    assert_rel(r["estimate"], case["estimate"], 1e-15)
    assert_rel(r["delta"], case["delta"], 1e-7)
    assert_rel(r["absvalue"], case["absvalue"], 1e-15)
    assert_rel(r["asc"], case["asc"], 1e-15)
    r = O.qk(case["rule"], case["functor"], case["b"], case["a"], case["params"])
    assert_rel(r["estimate"], -case["estimate"], 1e-15)
    assert_rel(r["delta"], case["delta"], 1e-7)
    assert_rel(r["absvalue"], case["absvalue"], 1e-15)
    assert_rel(r["asc"], case["asc"], 1e-15)


@pytest.mark.parametrize("case", G["single"], ids=lambda c: c["name"])
def test_single_driver(case):
    kw = dict(params=case["params"], points=case["points"], tol=tuple(case["tol"]),
              max_evals=case["max_evals"], ws_capacity=case["ws_capacity"])
    r = O.integrate_one(ALGO[case["algo"]], case["functor"], case["a"], case["b"], **kw)
    assert r["status"] == STATUS[case["status"]]
    assert_rel(r["estimate"], case["estimate"], 1e-15)
    assert_rel(r["delta"], case["delta"], 1e-7)
    assert r["nevals"] == case["nevals"]
    if case["order"]:
        assert r["order"] == case["order"]
    # reversed range (algorithms_single.rs:70-80): estimate negated; `order` only checked w/o points
    r = O.integrate_one(ALGO[case["algo"]], case["functor"], case["b"], case["a"], **kw)
    assert r["status"] == STATUS[case["status"]]
    assert_rel(r["estimate"], -case["estimate"], 1e-15)
    assert_rel(r["delta"], case["delta"], 1e-7)
    assert r["nevals"] == case["nevals"]
    if case["order"] and not case["points"]:
        assert r["order"] == case["order"]


def _run2(case, algo=None):
    return O.integrate2_batch(
        ALGO[algo or case["algo"]], case["functor2"], case["yrange"], num(case["x0"]), num(case["x1"]),
        fparams=case["fparams"] or None, yparams=[num(v) for v in case["yparams"]] or None,
        points=[tuple(p) for p in case["points"]], tol=tuple(case["tol"]),
        max_evals=case["max_evals"], n=1)


@pytest.mark.parametrize("case", G["double"], ids=lambda c: c["name"])
def test_double_driver(case):
    r = _run2(case)
    assert int(r["status"][0]) == STATUS[case["status"]]
    assert_rel(float(r["estimate"][0]), case["estimate"], 1e-15)
    assert_rel(float(r["delta"][0]), case["delta"], 1e-7)
    assert int(r["nevals"][0]) == case["nevals"]


@pytest.mark.parametrize("case", G["bench_asserts"], ids=lambda c: c["name"])
def test_bench_assertions(case):
    """benches/single.rs:15,27,37 and benches/double.rs:14,26,39 assert |result - truth| <= 1.49e-8;
    the probe table pins the full (estimate, delta, nevals, status) tuple as a regression value."""
    if case["dim"] == 1:
        r = O.integrate_one(ALGO[case["algo"]], case["functor"], num(case["a"]), num(case["b"]),
                            params=case["params"], points=case["points"], tol=tuple(case["tol"]),
                            max_evals=case["max_evals"])
        est, dlt, nev, st = r["estimate"], r["delta"], r["nevals"], r["status"]
    else:
        r = _run2(case)
        est, dlt, nev, st = (float(r["estimate"][0]), float(r["delta"][0]), int(r["nevals"][0]),
                             int(r["status"][0]))
    assert st == 0
    assert abs(est - case["truth"]) <= case["abs_bound"]
    probe = {c["name"]: c for c in G["probe_known_answers"]["cases"]}[case["name"]]
    assert_rel(est, probe["estimate"], 1e-15)
    assert_rel(dlt, probe["delta"], 1e-7)
    assert nev == probe["nevals"]


# ---- edge cases of SURVEY.md appendix A (reference behaviour by file:line given there) ----------
def test_edge_zero_width_range():
    r = O.integrate_one(O.ALGO_QAGS, "square", 0.5, 0.5)
    assert (r["estimate"], r["delta"], r["nevals"], r["status"]) == (0.0, 0.0, 17, 0)


def test_edge_max_evals_below_17():
    r = O.integrate_one(O.ALGO_QAGS, "f1", 0.0, 1.0, max_evals=16)
    assert (r["estimate"], r["nevals"], r["status"]) == (0.0, 0, 1)
    assert r["delta"] == 1.7976931348623157e308


def test_edge_max_evals_between_17_and_42():
    r = O.integrate_one(O.ALGO_QAGS, "f1", 0.0, 1.0, tol=("rel", 1e-10), max_evals=41)
    assert (r["nevals"], r["status"]) == (17, 1)


def test_edge_budget_skips_qk25_without_flag():
    # qags.rs:370-372,118,274-276: delta > 1024 tol after qk17 -> break; zero loop iterations -> None
    r = O.integrate_one(O.ALGO_QAGS, "f2", 0.0, 1.0, tol=("abs", 1e-10), max_evals=60)
    assert (r["nevals"], r["status"]) == (17, 0)


def test_edge_qags_budget_exhausted_reports_none():
    # the InsufficientIteration test at qags.rs:200 is unreachable
    r = O.integrate_one(O.ALGO_QAGS, "cos_p", 0.0, 1.0, params=[3000.0], tol=("rel", 1e-10), max_evals=2000)
    assert r["status"] == 0 and r["nevals"] in (1967, 1992) and r["delta"] > 1e-3


def test_edge_nan_integrand():
    r = O.integrate_one(O.ALGO_QAGS, "sqrt_shift_p", 0.0, 1.0, params=[0.5])
    assert r["status"] == 5 and math.isnan(r["estimate"]) and r["nevals"] == 17


def test_edge_infinite_value_is_not_nan():
    # 1/|x-1/2| hits +inf at the centre node: no flag, NaN delta, runs to the budget
    r = O.integrate_one(O.ALGO_QAGS, "inv_abs_shift_p", 0.0, 1.0, params=[0.5])
    assert r["status"] == 0 and math.isnan(r["delta"]) and r["nevals"] == 1992


def test_edge_half_infinite():
    r1 = O.integrate_one(O.ALGO_QAGS, "exp_p", 0.0, math.inf, params=[-1.0], tol=("rel", 1e-10))
    r2 = O.integrate_one(O.ALGO_QAGS, "exp_p", -math.inf, 0.0, params=[1.0], tol=("rel", 1e-10))
    assert r1["nevals"] == r2["nevals"] == 167 and r1["status"] == r2["status"] == 0
    assert r1["estimate"] == r2["estimate"] == 0.9999999999999999


def test_edge_qagp_points():
    base = O.integrate_one(O.ALGO_QAGP, "recip_p", -1.0, 1.0, params=[1.0], points=[0.0], tol=("abs", 1e-10))
    # points outside [min, max] are dropped (qagp.rs:397)
    r = O.integrate_one(O.ALGO_QAGP, "recip_p", -1.0, 1.0, params=[1.0], points=[0.0, 5.0, -7.0], tol=("abs", 1e-10))
    assert (r["estimate"], r["nevals"], r["status"]) == (base["estimate"], base["nevals"], base["status"])
    # unsorted points are sorted in range direction (qagp.rs:391-395)
    a = O.integrate_one(O.ALGO_QAGP, "f5", 0.0, 4.0, points=[2.0, 1.0], tol=("rel", 1e-3), ws_capacity=50)
    b = O.integrate_one(O.ALGO_QAGP, "f5", 0.0, 4.0, points=[1.0, 2.0], tol=("rel", 1e-3), ws_capacity=50)
    assert a == b
    # budget below 25*nint (qagp.rs:79-81)
    r = O.integrate_one(O.ALGO_QAGP, "f5", 0.0, 4.0, points=[1.0, 2.0], max_evals=74)
    assert (r["estimate"], r["nevals"], r["status"]) == (0.0, 0, 1)
    # budget == 25*nint and not converged (qagp.rs:160-166)
    r = O.integrate_one(O.ALGO_QAGP, "f5", 0.0, 4.0, points=[1.0, 2.0], tol=("rel", 1e-10), max_evals=75)
    assert (r["nevals"], r["status"]) == (75, 1)


def test_batch_matches_single_calls():
    n = 257
    rng = np.random.default_rng(1)
    p = np.stack([0.5 + 1.5 * rng.random(n), 1 + 9 * rng.random(n)])
    r = O.integrate_batch(O.ALGO_QAGS, "expcos_p", 0.0, 1.0, params=p, tol=("rel", 1e-10), n=n, nthreads=3)
    for i in (0, 1, 100, 256):
        s = O.integrate_one(O.ALGO_QAGS, "expcos_p", 0.0, 1.0, params=p[:, i], tol=("rel", 1e-10))
        assert s["estimate"] == r["estimate"][i] and s["nevals"] == r["nevals"][i]
    p0, p1 = p
    truth = (p0 * (1 - np.exp(-p0) * np.cos(p1)) + p1 * np.exp(-p0) * np.sin(p1)) / (p0 ** 2 + p1 ** 2)
    assert np.all(np.abs(r["estimate"] - truth) <= 1e-10 * np.abs(truth) + 1e-15)
    assert np.all(r["status"] == 0)


def test_ulp_probe_functor_stays_within_tolerance():
    """`expcos_p_ulp` (oracle-only: S1's integrand with cos() moved one ulp up, the probe behind
    tools/libm_sensitivity.py) must agree with `expcos_p` far inside the requested tolerance wherever
    both converge -- what changes under a 1-ulp libm difference is the occasional status / nevals."""
    import numpy as np
    from oracle import oracle as O
    n = 20000
    rng = np.random.default_rng(7)
    p = np.stack([0.5 + 1.5 * rng.random(n), 1.0 + 9.0 * rng.random(n)])
    kw = dict(params=p, tol=("rel", 1e-10), max_evals=2000, n=n)
    r0 = O.integrate_batch(O.ALGO_QAGS, "expcos_p", 0.0, 1.0, **kw)
    r1 = O.integrate_batch(O.ALGO_QAGS, "expcos_p_ulp", 0.0, 1.0, **kw)
    ok = (r0["status"] == 0) & (r1["status"] == 0)
    assert ok.sum() >= n - 20
    assert np.all(np.abs(r0["estimate"] - r1["estimate"])[ok] <= 1e-10 * np.abs(r0["estimate"])[ok])
    assert (r0["nevals"] != r1["nevals"]).sum() <= 5
