"""GPU parity tests: the CUDA engine, called through the C ABI (libgkquad_b200.so), against the
CPU oracle on the same seeded inputs and against the reference's golden vectors.

Bars (BASELINE.json:north_star / SURVEY.md section 7):
  * IEEE-exact integrands (+,-,*,/,sqrt): BIT-IDENTICAL estimate, delta, nevals, status;
  * transcendental integrands (libdevice vs glibc differ by ulps): |delta| <= max(abs_tol,
    rel_tol*|I|), identical status; nevals mismatches counted and bounded.
Run with `pytest -m gpu` on a B200.
"""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from helpers import ALGO, STATUS, assert_rel, num

pytestmark = pytest.mark.gpu

G = json.load(open(Path(__file__).parent / "golden" / "reference_vectors.json"))


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gkquad_b200 as gk
    return gk.default_engine(0)


def _tol(t):
    from gkquad_b200 import Tolerance
    k = t[0]
    return {"abs": lambda: Tolerance.Absolute(t[1]), "rel": lambda: Tolerance.Relative(t[1]),
            "abs_or_rel": lambda: Tolerance.AbsOrRel(t[1], t[2]),
            "abs_and_rel": lambda: Tolerance.AbsAndRel(t[1], t[2])}[k]()


# ---- golden vectors of the reference, through the GPU ----------------------------------------
# f1..f5 call libm (pow/ln/sin/cos): the reference's 1e-15 bound on the estimate is relaxed to
# 1e-13 for those (libdevice differs from glibc by <= 2 ulp per call); f6/f7 are IEEE-exact and
# keep 1e-15.  delta (1e-7), nevals and status stay exact.
_EXACT_FUNCTORS = {"f6", "recip_p", "square", "lorentz_p"}


@pytest.mark.parametrize("case", G["qk"], ids=lambda c: c["name"])
def test_golden_qk(eng, case):
    for a, b, sign in ((case["a"], case["b"], 1.0), (case["b"], case["a"], -1.0)):
        out = eng.qk_batch(case["rule"], case["functor"], [a], [b], case["params"] or None).cpu().numpy()[:, 0]
        assert_rel(out[0], sign * case["estimate"], 1e-13)
        # delta = rescaled (kronrod - gauss) amplifies ulp noise of the samples; when it sits at the
        # round-off floor (qk25_f3: 1.8e-14 on a result of 0.72) a few ulps move it by 1e-4 relative
        assert_rel(out[1], case["delta"], 1e-6 if case["delta"] > 1e-10 else 1e-3)
        assert_rel(out[2], case["absvalue"], 1e-13)
        assert_rel(out[3], case["asc"], 1e-13)


@pytest.mark.parametrize("rule", [17, 25, 33, 41, 49, 57])
def test_qk_batch_bit_exact(eng, rule):
    """Rule level (single/qk/mod.rs:28-73) on a batch of IEEE-exact integrands: every field of QKResult
    bit-identical to the oracle, forward and reversed ranges."""
    from oracle import oracle as O
    rng = np.random.default_rng(rule)
    n = 257
    for f, lo, hi in (("lorentz_p", -3.0, 3.0), ("recip_p", 0.1, 5.0)):
        p = rng.uniform(0.5, 50.0, (1, n))
        a = rng.uniform(lo, 0.5 * (lo + hi), n)
        b = rng.uniform(0.5 * (lo + hi), hi, n)
        a[::3], b[::3] = b[::3].copy(), a[::3].copy()  # reversed ranges
        out = eng.qk_batch(rule, f, a, b, p).cpu().numpy()
        for i in range(n):
            r = O.qk(rule, f, float(a[i]), float(b[i]), [float(p[0, i])])
            got = out[:, i]
            assert (got[0], got[1], got[2], got[3]) == (r["estimate"], r["delta"], r["absvalue"], r["asc"]), (f, i)


@pytest.mark.parametrize("case", G["single"], ids=lambda c: c["name"])
def test_golden_single(eng, case):
    etol = 1e-15 if case["functor"] in _EXACT_FUNCTORS else 1e-13
    for a, b, sign in ((case["a"], case["b"], 1.0), (case["b"], case["a"], -1.0)):
        r = eng.integrate_batch(case["functor"], a, b, case["params"] or None, algo=ALGO[case["algo"]],
                                tol=_tol(case["tol"]), max_evals=case["max_evals"], points=case["points"],
                                workspace_capacity=case["ws_capacity"], n=1)[0]
        assert (r.error.value if r.error else 0) == STATUS[case["status"]], r
        assert_rel(r.value.estimate, sign * case["estimate"], etol)
        if case["name"] != "qags_f2":  # knife-edge case: SURVEY section 7 (nevals scatters under ulp noise)
            assert r.value.nevals == case["nevals"], r
            # a delta at the round-off floor moves by ~1e-4 relative under ulp-level sample noise
            assert_rel(r.value.delta, case["delta"], 1e-6 if case["delta"] > 1e-10 * abs(case["estimate"]) else 1e-3)


@pytest.mark.parametrize("case", G["double"], ids=lambda c: c["name"])
def test_golden_double(eng, case):
    r = eng.integrate2_batch(case["functor2"], case["yrange"], num(case["x0"]), num(case["x1"]),
                             fparams=case["fparams"] or None,
                             yparams=[num(v) for v in case["yparams"]] or None, algo=ALGO[case["algo"]],
                             tol=_tol(case["tol"]), max_evals=case["max_evals"],
                             points=[tuple(p) for p in case["points"]], n=1)[0]
    assert (r.error.value if r.error else 0) == STATUS[case["status"]], r
    exact = case["functor2"] != "g2"  # g2 = exp(y/x) is the only transcendental one
    assert_rel(r.value.estimate, case["estimate"], 1e-15 if exact else 1e-13)
    assert_rel(r.value.delta, case["delta"], 1e-7 if exact else 1e-5)
    assert r.value.nevals == case["nevals"], r


@pytest.mark.parametrize("case", G["bench_asserts"], ids=lambda c: c["name"])
def test_bench_assertions(eng, case):
    probe = {c["name"]: c for c in G["probe_known_answers"]["cases"]}[case["name"]]
    if case["dim"] == 1:
        r = eng.integrate_batch(case["functor"], num(case["a"]), num(case["b"]), case["params"] or None,
                                algo=ALGO[case["algo"]], tol=_tol(case["tol"]), max_evals=case["max_evals"],
                                points=case["points"], n=1)[0]
    else:
        r = eng.integrate2_batch(case["functor2"], case["yrange"], num(case["x0"]), num(case["x1"]),
                                 fparams=case["fparams"] or None, yparams=[num(v) for v in case["yparams"]] or None,
                                 algo=ALGO[case["algo"]], tol=_tol(case["tol"]), max_evals=case["max_evals"],
                                 points=[tuple(p) for p in case["points"]], n=1)[0]
    assert r.error is None
    assert abs(r.value.estimate - case["truth"]) <= case["abs_bound"]
    # every bench integrand is IEEE-exact: bit-equal to the oracle's regenerated values
    assert r.value.estimate == probe["estimate"] and r.value.nevals == probe["nevals"], r
    assert_rel(r.value.delta, probe["delta"], 1e-12)


# ---- seeded batches against the oracle ---------------------------------------------------------
def _run_both(eng, wname, n, start=0, device=True):
    import torch
    from gkquad_b200.workloads import WORKLOADS, csr_points_from_param0
    from gpu_common import compare, tol_to_oracle
    from oracle import oracle as O
    W = WORKLOADS[wname]
    p = W.make_params(start, n)
    kw_g, kw_o = {}, {}
    if W.points == "param0":
        off, pts = csr_points_from_param0(p)
        kw_g = dict(pts_offsets=off, pts=pts)
        kw_o = dict(pts_offsets=off, points=pts)
    elif W.points:
        kw_g = dict(points=W.points)
        kw_o = dict(points=W.points)
    pg = torch.as_tensor(p, device="cuda:0") if device else p
    g = eng.integrate_batch(W.functor, W.a, W.b, pg, algo=ALGO[W.algo], tol=W.tol, max_evals=W.max_evals,
                            n=n, **kw_g)
    ref = O.integrate_batch(ALGO[W.algo], W.functor, W.a, W.b, params=p, tol=tol_to_oracle(W.tol),
                            max_evals=W.max_evals, n=n, **kw_o)
    return W, p, g, ref


@pytest.mark.parametrize("wname", ["S2", "P0", "I0", "I1"])
def test_batch_bit_exact(eng, wname):
    """IEEE-exact integrands: estimate, delta, nevals and status bit-identical to the oracle."""
    from gpu_common import compare
    W, p, g, ref = _run_both(eng, wname, 20000)
    s = compare(g, ref, W.tol, exact=True, label=wname)
    if W.truth is not None:
        truth = W.truth(p)
        ok = ref["status"] == 0
        from gpu_common import allowed_error
        assert np.all(np.abs(g.numpy().estimate - truth)[ok] <= allowed_error(W.tol, truth)[ok] + 1e-14)


@pytest.mark.parametrize("wname,max_nevals_mismatch", [("S1", 0.002), ("S3", 0.01), ("I2", 0.01),
                                                        ("P1", 0.10), ("P2", 0.10)])
def test_batch_tolerance_parity(eng, wname, max_nevals_mismatch):
    """Transcendental integrands: within tolerance of the oracle, identical status -- except for
    noise-level flips (see gpu_common.compare): at most 1 in these 20 000-integral samples, 2 for the
    stress family S3 (measured: 4 per 2^20 on S1, 33 on S3; the oracle flips 7 per 10^6 against itself
    when its cos moves by one ulp, tools/libm_sensitivity.py), every one of them with both values
    agreeing within the tolerance."""
    from gpu_common import compare
    n = 20000 if wname.startswith("S") else 4000
    W, p, g, ref = _run_both(eng, wname, n)
    s = compare(g, ref, W.tol, label=wname)
    print(s)
    assert s["out_of_tol"] == 0, s
    # (measured floor: 4 flips per 2^20 on S1, 33 on S3 -> expect 0.08 resp. 0.6 in 20 000; the full-size
    # bound is asserted by test_gpu_round2.py::test_full_size_tolerance_parity)
    assert s["status_mismatch"] == s["noise_level_flips"] <= (2 if wname == "S3" else 1), s
    assert s["nevals_mismatch"] <= max_nevals_mismatch * n, s


def test_host_entry_matches_device_entry(eng):
    """gkq_integrate_batch_host (chunked H2D/compute/D2H pipeline) == device entry, bit for bit."""
    from gkquad_b200.workloads import WORKLOADS
    n = 300_001  # > one 2^18 chunk, ragged tail
    W, p, g_dev, _ = (None, None, None, None)
    import torch
    W = WORKLOADS["S1"]
    p = W.make_params(0, n)
    g_host = eng.integrate_batch(W.functor, W.a, W.b, p, algo=ALGO[W.algo], tol=W.tol, n=n)
    g_dev = eng.integrate_batch(W.functor, W.a, W.b, torch.as_tensor(p, device="cuda:0"), algo=ALGO[W.algo],
                                tol=W.tol, n=n).numpy()
    assert np.array_equal(g_host.estimate, g_dev.estimate) and np.array_equal(g_host.delta, g_dev.delta)
    assert np.array_equal(g_host.nevals, g_dev.nevals) and np.array_equal(g_host.status, g_dev.status)
    # per-integral ranges through the host entry
    a = np.zeros(n)
    b = np.ones(n)
    g2 = eng.integrate_batch(W.functor, a, b, p, algo=ALGO[W.algo], tol=W.tol)
    assert np.array_equal(g2.estimate, g_dev.estimate)


@pytest.mark.parametrize("wname", ["I1", "S3", "I2"])
def test_host_entry_heavy_tail(eng, wname):
    """Host entry when a large share of the batch needs the persistent kernel (every integral of an
    infinite-range batch takes more than one bisection step): the remainder writes into the staged
    arrays, which are copied out whole, instead of a million-entry patch applied by the CPU.  Bit for
    bit the device entry's results; also across the threshold between the two paths."""
    import torch
    from gkquad_b200.workloads import WORKLOADS
    W = WORKLOADS[wname]
    for n in (70_001, 300_001):
        p = W.make_params(0, n)
        g_host = eng.integrate_batch(W.functor, W.a, W.b, p, algo=ALGO[W.algo], tol=W.tol, n=n)
        g_dev = eng.integrate_batch(W.functor, W.a, W.b, torch.as_tensor(p, device="cuda:0"), algo=ALGO[W.algo],
                                    tol=W.tol, n=n).numpy()
        assert np.array_equal(g_host.estimate, g_dev.estimate) and np.array_equal(g_host.delta, g_dev.delta)
        assert np.array_equal(g_host.nevals, g_dev.nevals) and np.array_equal(g_host.status, g_dev.status)
        # again: the second call reuses the other record buffer and the staging area
        g_host2 = eng.integrate_batch(W.functor, W.a, W.b, p, algo=ALGO[W.algo], tol=W.tol, n=n)
        assert np.array_equal(g_host2.estimate, g_dev.estimate) and np.array_equal(g_host2.nevals, g_dev.nevals)


def test_full_size_properties(eng):
    """BASELINE cfg-2 size (1M, S1): properties that need no oracle -- closed-form truth within
    tolerance, reversed range negates exactly, results independent of batch position."""
    import torch
    from gkquad_b200.workloads import WORKLOADS
    W = WORKLOADS["S1"]
    n = 1 << 20
    p = W.make_params(0, n)
    pt = torch.as_tensor(p, device="cuda:0")
    g = eng.integrate_batch(W.functor, 0.0, 1.0, pt, algo=ALGO["QAGS"], tol=W.tol, n=n).numpy()
    truth = W.truth(p)
    # a few 1e-4 of S1 stop with RoundoffError at the initial rule (the oracle does the same)
    assert set(np.unique(g.status)).issubset({0, 2}) and (g.status == 0).mean() > 0.999
    ok = g.status == 0
    assert np.all(np.abs(g.estimate - truth)[ok] <= 1e-10 * np.abs(truth)[ok] + 1e-16)
    assert np.all(np.abs(g.estimate - truth) <= 1e-9 * np.abs(truth) + 1e-15)
    assert np.all((g.nevals - 17) % 25 == 0) and g.nevals.max() <= 1992
    rev = eng.integrate_batch(W.functor, 1.0, 0.0, pt, algo=ALGO["QAGS"], tol=W.tol, n=n).numpy()
    assert np.array_equal(rev.estimate, -g.estimate) and np.array_equal(rev.nevals, g.nevals)
    # a shuffled batch gives the same per-integral answers (no cross-lane leakage)
    perm = np.random.default_rng(0).permutation(n)
    gs = eng.integrate_batch(W.functor, 0.0, 1.0, torch.as_tensor(np.ascontiguousarray(p[:, perm]), device="cuda:0"),
                             algo=ALGO["QAGS"], tol=W.tol, n=n).numpy()
    assert np.array_equal(gs.estimate, g.estimate[perm]) and np.array_equal(gs.nevals, g.nevals[perm])


def test_large_batch_16m(eng):
    """A batch 16x the headline size through both entries (worklists, straggler records, host chunks and
    slabs all scale with n): IEEE-exact integrand, so the halves of the batch -- same parameters -- must
    be bit-identical to each other and to the oracle on a 1M slice."""
    import torch
    from gkquad_b200 import Tolerance
    from gpu_common import compare
    from oracle import oracle as O
    m = 1 << 20
    n = 16 * m
    rng = np.random.default_rng(1)
    base = 10.0 ** rng.uniform(0, 3.5, m)
    p = np.ascontiguousarray(np.tile(base, 16)[None, :])
    tol = Tolerance.Relative(1e-10)
    ref = O.integrate_batch(ALGO["QAGS"], "lorentz_p", -1.0, 1.0, p[:, :m], tol=("rel", 1e-10), max_evals=2000, n=m)
    g = eng.integrate_batch("lorentz_p", -1.0, 1.0, torch.as_tensor(p, device="cuda:0"), algo=ALGO["QAGS"], tol=tol,
                            n=n).numpy()
    for k in range(16):
        s = slice(k * m, (k + 1) * m)
        assert np.array_equal(g.estimate[s], ref["estimate"]) and np.array_equal(g.nevals[s], ref["nevals"])
        assert np.array_equal(g.status[s], ref["status"]) and np.array_equal(g.delta[s], ref["delta"])
    h = eng.integrate_batch("lorentz_p", -1.0, 1.0, p, algo=ALGO["QAGS"], tol=tol, n=n).numpy()
    assert np.array_equal(h.estimate, g.estimate) and np.array_equal(h.delta, g.delta)
    assert np.array_equal(h.nevals, g.nevals) and np.array_equal(h.status, g.status)
    assert len(np.unique(ref["nevals"])) > 4 and (ref["nevals"] > 67).sum() > 1000  # the adaptive tail was exercised


# ---- edge cases against the oracle ------------------------------------------
EDGE = [
    dict(f="square", a=0.5, b=0.5, algo="QAGS"),
    dict(f="f1", a=0.0, b=1.0, algo="QAGS", max_evals=16),
    dict(f="lorentz_p", p=[3.0], a=0.0, b=1.0, algo="QAGS", tol=("rel", 1e-10), max_evals=41),
    dict(f="lorentz_p", p=[3.0], a=0.0, b=1.0, algo="QAGS", tol=("rel", 1e-10), max_evals=66),
    dict(f="recip_p", p=[1.0], a=0.0, b=1.0, algo="QAGS", tol=("abs", 1e-10), max_evals=60),
    dict(f="sqrt_shift_p", p=[0.5], a=0.0, b=1.0, algo="QAGS"),
    dict(f="sqrt_shift_p", p=[0.3], a=0.0, b=1.0, algo="QAGP", points=[0.5]),
    dict(f="inv_abs_shift_p", p=[0.5], a=0.0, b=1.0, algo="QAGS"),
    dict(f="lorentz_p", p=[2.0], a=0.0, b=math.inf, algo="QAGS", tol=("rel", 1e-10)),
    dict(f="lorentz_p", p=[2.0], a=-math.inf, b=0.0, algo="QAGS", tol=("rel", 1e-10)),
    dict(f="lorentz_p", p=[2.0], a=math.inf, b=-math.inf, algo="QAGS", tol=("rel", 1e-10)),
    dict(f="recip_p", p=[1.0], a=-1.0, b=1.0, algo="QAGP", points=[0.0, 5.0, -7.0], tol=("abs", 1e-10)),
    dict(f="recip_p", p=[1.0], a=-1.0, b=1.0, algo="QAGP", points=[0.0, 0.0, 1.0], tol=("abs", 1e-10)),
    dict(f="f6", a=0.0, b=2.0, algo="QAGP", points=[1.0], tol=("abs", 1e-10)),
    dict(f="f6", a=2.0, b=0.0, algo="QAGP", points=[1.0, 0.5, 1.5], tol=("abs", 1e-10)),
    dict(f="f6", a=0.0, b=2.0, algo="QAGP", points=[1.0, 0.5], max_evals=74),
    dict(f="f6", a=0.0, b=2.0, algo="QAGP", points=[1.0, 0.5], tol=("abs", 1e-12), max_evals=75),
    dict(f="f6", a=0.0, b=2.0, algo="QAGP", points=[1.0], tol=("abs", 1e-12), max_evals=130),
    dict(f="f6", a=0.0, b=2.0, algo="AUTO", points=[], tol=("abs", 1e-6)),
    dict(f="lorentz_p", p=[1e4], a=-1.0, b=1.0, algo="QAGS", tol=("rel", 1e-13)),
    dict(f="lorentz_p", p=[1e4], a=-1.0, b=1.0, algo="QAGS", tol=("rel", 1e-13), max_evals=400),
    dict(f="recip_p", p=[1.0], a=0.0, b=1.0, algo="QAGS", tol=("rel", 1e-8)),          # divergent 1/x
    dict(f="inv_abs_shift_p", p=[0.3], a=0.0, b=1.0, algo="QAGP", points=[0.3], tol=("rel", 1e-8)),
    dict(f="lorentz_p", p=[1.0], a=-math.inf, b=math.inf, algo="QAGP", points=[0.0, 1.0], tol=("rel", 1e-10)),
    dict(f="lorentz_p", p=[1.0], a=0.0, b=1.0, algo="QAGS", tol=("abs_and_rel", 1e-6, 1e-9)),
    dict(f="lorentz_p", p=[1.0], a=0.0, b=1.0, algo="QAGS", tol=("abs_or_rel", 1e-12, 1e-4)),
]


@pytest.mark.parametrize("c", EDGE, ids=lambda c: f"{c['f']}-{c['algo']}-{c.get('max_evals', 2000)}-{c['a']}")
def test_edge_cases_bit_exact(eng, c):
    """All edge-case integrands are IEEE-exact -> the GPU must reproduce the oracle bit for bit
    (estimate, delta, nevals, status), including NaN/inf poisoning and the status quirks."""
    from oracle import oracle as O
    tol = c.get("tol", ("abs_or_rel", 1.49e-8, 1.49e-8))
    me = c.get("max_evals", 2000)
    ref = O.integrate_one(ALGO[c["algo"]], c["f"], c["a"], c["b"], params=c.get("p", []), points=c.get("points", []),
                          tol=tol, max_evals=me)
    r = eng.integrate_batch(c["f"], c["a"], c["b"], c.get("p"), algo=ALGO[c["algo"]], tol=_tol(tol), max_evals=me,
                            points=c.get("points", []), n=1)[0]
    st = r.error.value if r.error else 0
    assert st == ref["status"], (r, ref)
    assert r.value.nevals == ref["nevals"], (r, ref)
    for got, want in ((r.value.estimate, ref["estimate"]), (r.value.delta, ref["delta"])):
        assert (math.isnan(got) and math.isnan(want)) or got == want, (r, ref)


def test_workspace_capacity_changes_limit(eng):
    """`limit` = Vec capacity (workspace.rs:165,236): a too-small persistent capacity changes the
    bisection order; the engine must follow the oracle for both."""
    from oracle import oracle as O
    for cap in (0, 50, 12):
        ref = O.integrate_one(O.ALGO_QAGS, "lorentz_p", -1.0, 1.0, params=[1e4], tol=("rel", 1e-13),
                              max_evals=500, ws_capacity=cap)
        r = eng.integrate_batch("lorentz_p", -1.0, 1.0, [1e4], algo=ALGO["QAGS"], tol=_tol(("rel", 1e-13)),
                                max_evals=500, workspace_capacity=cap, n=1)[0]
        assert (r.value.estimate, r.value.delta, r.value.nevals) == (ref["estimate"], ref["delta"], ref["nevals"])


def test_mixed_auto_csr(eng):
    """AUTO over per-integral points: integrals without points run QAGS, the others QAGP
    (auto.rs:29-33); ragged CSR incl. empty rows, duplicated and out-of-range points."""
    import torch
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    n = 5000
    cnt = rng.integers(0, 4, n)
    cnt[:7] = [0, 1, 0, 3, 2, 0, 1]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    pts = rng.uniform(-1.2, 1.2, off[-1])
    p = (0.5 + 50 * rng.random(n))[None, :]
    ref = O.integrate_batch(O.ALGO_AUTO, "lorentz_p", -1.0, 1.0, params=p, points=pts, pts_offsets=off,
                            tol=("rel", 1e-12), n=n)
    g = eng.integrate_batch("lorentz_p", -1.0, 1.0, torch.as_tensor(p, device="cuda:0"), algo=ALGO["AUTO"],
                            tol=_tol(("rel", 1e-12)), pts_offsets=off, pts=pts, n=n)
    from gpu_common import compare
    compare(g, ref, _tol(("rel", 1e-12)), exact=True)


# ---- 2-D batches ---------------------------------------------------------------------------------
def test_double_batch_bit_exact(eng):
    """cfg-5 shapes at oracle-friendly sizes: D0 (literal bench, parameterised) and D1 (singular point
    per integral, CSR).  All 2-D bench integrands are IEEE-exact -> bit-identical to the oracle."""
    import torch
    from gkquad_b200.workloads import WORKLOADS
    from gpu_common import compare, tol_to_oracle
    from oracle import oracle as O
    W = WORKLOADS["D0"]
    n = 512
    p = W.make_params(0, n)
    ref = O.integrate2_batch(ALGO[W.algo], W.functor, W.yrange, W.a, W.b, fparams=p, yparams=W.yparams,
                             points=W.points, tol=tol_to_oracle(W.tol), max_evals=W.max_evals, n=n)
    g = eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=torch.as_tensor(p, device="cuda:0"),
                             yparams=W.yparams, algo=ALGO[W.algo], tol=W.tol, max_evals=W.max_evals,
                             points=W.points, n=n)
    compare(g, ref, W.tol, exact=True, label="D0")
    assert np.all(g.numpy().nevals == 12500) and np.all(g.numpy().estimate == 0.0)

    W = WORKLOADS["D1"]
    n = 96
    p = W.make_params(0, n)
    off = np.arange(n + 1, dtype=np.int64)
    pxy = np.ascontiguousarray(p.T)
    ref = O.integrate2_batch(ALGO[W.algo], W.functor, W.yrange, W.a, W.b, fparams=p, yparams=W.yparams,
                             points=pxy, pts_offsets=off, tol=tol_to_oracle(W.tol), max_evals=W.max_evals, n=n)
    g = eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=p, yparams=W.yparams, algo=ALGO[W.algo],
                             tol=W.tol, max_evals=W.max_evals, pts_offsets=off, pts_xy=pxy, n=n)
    s = compare(g, ref, W.tol, exact=True, label="D1")
    print(s, np.bincount(ref["status"]))


def test_double_dynamic_ranges(eng):
    """DynamicY ranges (0..x, circle) and infinite x infinite rectangles against the oracle."""
    from oracle import oracle as O
    from gpu_common import compare
    cases = [("g1", "const", 0.0, 1.0, [0.0, 1.0], ("rel", 1e-10), ALGO["QAGS"], []),
             ("g3", "circle", -0.5, 0.5, [0.25], ("abs", 1e-12), ALGO["QAGS"], []),
             ("g4", "const", 0.0, math.inf, [0.0, math.inf], ("abs", 1e-8), ALGO["QAGS"], []),
             ("rsqrt3", "const", -math.inf, math.inf, [-math.inf, math.inf], ("abs_or_rel", 1.49e-8, 1.49e-8), 0, []),
             ("gp2", "circle", -1.0, 1.0, [1.0], ("abs", 1e-5), ALGO["QAGP"], [(0.0, 0.0)]),
             ("gp1", "const", -1.0, math.inf, [-math.inf, 2.0], ("rel", 1e-4), ALGO["QAGP"], [(0.0, 0.0), (1.0, 1.0)])]
    for f, yr, x0, x1, yq, tol, algo, pts in cases:
        ref = O.integrate2_batch(algo, f, yr, x0, x1, yparams=yq, points=pts, tol=tol, max_evals=100000, n=1)
        g = eng.integrate2_batch(f, yr, x0, x1, yparams=yq, algo=algo, tol=_tol(tol), max_evals=100000,
                                 points=pts, n=1)
        compare(g, ref, _tol(tol), exact=True, label=f)


@pytest.mark.parametrize("mode", ["1", "2", "3"], ids=["thread_per_outer", "warp_per_outer", "level_synchronous"])
def test_double_budget_exhaustion(eng, mode, monkeypatch):
    """The running evaluation budget, the first-error latch and the persistent inner workspace of the
    nest (double/algorithm/qags.rs:60-93, qagp.rs:108-137) under budgets that run out at every stage;
    both device mappings (GKQ_NESTED_MODE) must agree bit for bit with the oracle."""
    import torch
    from gkquad_b200.workloads import WORKLOADS
    from gpu_common import compare
    from oracle import oracle as O
    monkeypatch.setenv("GKQ_NESTED_MODE", mode)
    W = WORKLOADS["D1"]
    n = 24
    p = W.make_params(0, n)
    off = np.arange(n + 1, dtype=np.int64)
    pxy = np.ascontiguousarray(p.T)
    seen = set()
    for me in (0, 16, 49, 50, 120, 700, 1250, 1300, 2600, 5000, 12000, 30000, 60000, 85000, 92000):
        for tol in (("rel", 1e-6), ("abs", 1e-9)):
            ref = O.integrate2_batch(ALGO["QAGP"], W.functor, W.yrange, W.a, W.b, fparams=p, yparams=W.yparams,
                                     points=pxy, pts_offsets=off, tol=tol, max_evals=me, n=n)
            g = eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=p, yparams=W.yparams,
                                     algo=ALGO["QAGP"], tol=_tol(tol), max_evals=me, pts_offsets=off, pts_xy=pxy, n=n)
            compare(g, ref, _tol(tol), exact=True, label=f"D1 max_evals={me}")
            seen |= set(np.unique(ref["status"]).tolist())
    assert 1 in seen  # InsufficientIteration was exercised
    cases = [("g1", "const", 0.0, 1.0, [0.0, 1.0], ALGO["QAGS"], []),
             ("g2", "zero_to_x", 0.1, 1.0, [], ALGO["QAGS"], []),
             ("gp1", "const", -1.0, 1.0, [-1.0, 1.0], ALGO["QAGS"], []),
             ("gp3", "const", -1.0, 2.0, [-1.0, 1.0], ALGO["QAGP"], [(0.0, 0.0)]),
             ("gp1", "const", -1.0, math.inf, [-math.inf, 2.0], ALGO["QAGP"], [(0.0, 0.0), (1.0, 1.0)])]
    for f, yr, x0, x1, yq, algo, pts in cases:
        for me in (10, 17, 288, 289, 290, 700, 1100, 3000, 9000, 25000):
            for tol in (("rel", 1e-9), ("abs", 1e-13)):
                ref = O.integrate2_batch(algo, f, yr, x0, x1, yparams=yq, points=pts, tol=tol, max_evals=me, n=1)
                g = eng.integrate2_batch(f, yr, x0, x1, yparams=yq, algo=algo, tol=_tol(tol), max_evals=me,
                                         points=pts, n=1)
                compare(g, ref, _tol(tol), exact=f != "g2", label=f"{f} max_evals={me}")


def test_qag_batch_bit_exact(eng):
    """The deprecated QAG (single/algorithm/qag.rs:65-248) on batches of IEEE-exact integrands: finite and
    infinite ranges, per-integral budgets, device and host entry -- bit-identical to the oracle."""
    import torch
    from gkquad_b200 import Tolerance
    from gpu_common import compare
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    n = 6000
    p = 10.0 ** rng.uniform(0, 4, (1, n))
    me = rng.choice([0, 16, 17, 41, 42, 66, 67, 117, 500, 2000], n).astype(np.uint64)
    for a, b, tol, otol in ((-1.0, 1.0, Tolerance.Relative(1e-11), ("rel", 1e-11)),
                            (-math.inf, math.inf, Tolerance.Absolute(1e-10), ("abs", 1e-10)),
                            (0.0, 1.0, Tolerance.Absolute(1e-15), ("abs", 1e-15))):
        ref = O.integrate_batch(ALGO["QAG"], "lorentz_p", a, b, p, tol=otol, max_evals=2000, n=n, max_evals_each=me)
        g = eng.integrate_batch("lorentz_p", a, b, torch.as_tensor(p, device="cuda:0"), algo=ALGO["QAG"], tol=tol,
                                max_evals=2000, n=n, max_evals_each=me.astype(np.int64))
        compare(g, ref, tol, exact=True, label=f"QAG device [{a}, {b}]")
        g = eng.integrate_batch("lorentz_p", a, b, p, algo=ALGO["QAG"], tol=tol, max_evals=2000, n=n, max_evals_each=me)
        compare(g, ref, tol, exact=True, label=f"QAG host [{a}, {b}]")
    assert len(np.unique(ref["status"])) >= 2


def test_per_integral_tolerance_and_budget(eng):
    """every integral with its own tolerance values and evaluation budget (an Integrator
    re-configured per integral, single/integrator.rs:69-84) -- device entry, host entry, QAGS and QAGP,
    bit-exact against the oracle on IEEE-exact integrands."""
    import torch
    from gkquad_b200 import Tolerance
    from gpu_common import compare
    from oracle import oracle as O
    rng = np.random.default_rng(7)
    n = 20000
    p = rng.uniform(1.0, 3000.0, (1, n))
    rel = 10.0 ** rng.uniform(-13, -3, n)
    ab = 10.0 ** rng.uniform(-14, -6, n)
    me = rng.choice([0, 10, 17, 41, 42, 66, 67, 92, 117, 500, 2000, 5000], n).astype(np.uint64)
    for kind, tol in (("rel", Tolerance.Relative(1e-9)), ("abs_or_rel", Tolerance.AbsOrRel(1e-12, 1e-9))):
        otol = ("rel", 1e-9) if kind == "rel" else ("abs_or_rel", 1e-12, 1e-9)
        ref = O.integrate_batch(ALGO["QAGS"], "lorentz_p", -1.0, 1.0, p, tol=otol, max_evals=2000, n=n,
                                abs_tol_each=ab, rel_tol_each=rel, max_evals_each=me)
        assert len(np.unique(ref["status"])) >= 2 and len(np.unique(ref["nevals"])) > 6
        # device entry
        g = eng.integrate_batch("lorentz_p", -1.0, 1.0, torch.as_tensor(p, device="cuda:0"), algo=ALGO["QAGS"],
                                tol=tol, max_evals=2000, n=n, abs_tol_each=ab, rel_tol_each=rel,
                                max_evals_each=me.astype(np.int64))
        compare(g, ref, tol, exact=True, label=f"device {kind}")
        # host entry
        g = eng.integrate_batch("lorentz_p", -1.0, 1.0, p, algo=ALGO["QAGS"], tol=tol, max_evals=2000, n=n,
                                abs_tol_each=ab, rel_tol_each=rel, max_evals_each=me)
        compare(g, ref, tol, exact=True, label=f"host {kind}")
    # QAGP with a shared break point
    pq = rng.uniform(0.5, 2.0, (1, n))
    ref = O.integrate_batch(ALGO["QAGP"], "recip_p", -1.0, 1.0, pq, points=[0.0], tol=("abs", 1e-8), max_evals=2000,
                            n=n, abs_tol_each=ab, max_evals_each=me)
    g = eng.integrate_batch("recip_p", -1.0, 1.0, torch.as_tensor(pq, device="cuda:0"), algo=ALGO["QAGP"],
                            tol=Tolerance.Absolute(1e-8), max_evals=2000, points=[0.0], n=n, abs_tol_each=ab,
                            max_evals_each=me.astype(np.int64))
    compare(g, ref, Tolerance.Absolute(1e-8), exact=True, label="QAGP")
    # the deferred-join pipeline keeps each record's own configuration
    eng.set_deferred_join(True)
    try:
        outs = []
        for k in range(3):
            outs.append(eng.integrate_batch("lorentz_p", -1.0, 1.0, torch.as_tensor(p, device="cuda:0"),
                                            algo=ALGO["QAGS"], tol=Tolerance.Relative(1e-9), max_evals=2000, n=n,
                                            rel_tol_each=rel * (10.0 ** -k), max_evals_each=me.astype(np.int64)))
        eng.join()
        torch.cuda.synchronize()
        for k, g in enumerate(outs):
            ref = O.integrate_batch(ALGO["QAGS"], "lorentz_p", -1.0, 1.0, p, tol=("rel", 1e-9), max_evals=2000, n=n,
                                    rel_tol_each=rel * (10.0 ** -k), max_evals_each=me)
            compare(g, ref, Tolerance.Relative(1e-9), exact=True, label=f"deferred {k}")
    finally:
        eng.set_deferred_join(False)


@pytest.mark.parametrize("mode", ["1", "2", "3"], ids=["thread_per_outer", "pool", "level_synchronous"])
def test_double_dynamic_x(eng, mode, monkeypatch):
    """DynamicX (double/range.rs:38-75; algorithms: qags.rs:29-44, qagp.rs:34-53): x inside, y outside,
    integrand arguments and point coordinates swapped.  Bit-exact against the oracle, and equal to the
    DynamicY result of the transposed problem where the integrand is symmetric."""
    from gkquad_b200 import Tolerance, double
    from gpu_common import compare
    from oracle import oracle as O
    monkeypatch.setenv("GKQ_NESTED_MODE", mode)
    cases = [("gp2", "circle", -1.0, 1.0, [1.0], ("abs", 1e-5), ALGO["QAGP"], [(0.0, 0.0)]),     # x / (x^2 + y^2)
             ("g2", "zero_to_x", 0.1, 1.0, [], ("rel", 1e-8), ALGO["QAGS"], []),                  # exp(y / x)
             ("g4", "const", 0.0, 2.0, [0.0, 3.0], ("abs", 1e-9), ALGO["QAGS"], []),
             ("gp1", "const", -1.0, 2.0, [-1.0, 1.0], ("rel", 1e-4), ALGO["QAGP"], [(0.5, 0.0), (0.0, 0.0)])]
    for f, xr, y1, y2, xq, tol, algo, pts in cases:
        pts_sw = [(y, x) for (x, y) in pts]
        ref = O.integrate2_batch(algo, f, xr, y1, y2, yparams=xq, points=pts_sw, tol=tol, max_evals=100000, n=1,
                                 swap_xy=True)
        g = eng.integrate2_batch(f, xr, y1, y2, yparams=xq, algo=algo, tol=_tol(tol), max_evals=100000,
                                 points=pts_sw, n=1, swap_xy=True)
        compare(g, ref, _tol(tol), exact=f != "g2", label=f"DynamicX {f}")
        plain = O.integrate2_batch(algo, f, xr, y1, y2, yparams=xq, points=pts_sw, tol=tol, max_evals=100000, n=1)
        if f in ("gp2", "g2"):  # not symmetric in (x, y): the swap must change the answer
            assert plain["estimate"][0] != ref["estimate"][0]
        else:                   # g4, gp1 are symmetric in (x, y) up to the rounding of 1 + x + y
            assert abs(plain["estimate"][0] - ref["estimate"][0]) <= 1e-13 * abs(ref["estimate"][0])
            assert plain["nevals"][0] == ref["nevals"][0]
    # through the mirror of the reference API
    it = double.Integrator2.new(double.DeviceIntegrand2("gp1")).tolerance(Tolerance.Relative(1e-4)).points([(0.0, 0.5)])
    r = it.run(double.DynamicX.new("const", -1.0, 2.0, [-1.0, 1.0]))
    ref = O.integrate2_batch(ALGO["QAGP"], "gp1", "const", -1.0, 2.0, yparams=[-1.0, 1.0], points=[(0.5, 0.0)],
                             tol=("rel", 1e-4), max_evals=100000, n=1, swap_xy=True)
    assert r.value.estimate == ref["estimate"][0] and r.value.nevals == ref["nevals"][0]


@pytest.mark.parametrize("mode", ["1", "2", "3"], ids=["thread_per_outer", "pool", "level_synchronous"])
def test_double_per_integral_tolerance_and_budget(eng, mode, monkeypatch):
    """Per-nest tolerance values and TOTAL budgets (IntegrationConfig2, double/common.rs:14-32, set anew
    for every integral of a sweep): bit-exact against the oracle, device and host entry."""
    import torch
    from gkquad_b200 import Tolerance
    from gkquad_b200.workloads import WORKLOADS
    from gpu_common import compare
    from oracle import oracle as O
    monkeypatch.setenv("GKQ_NESTED_MODE", mode)
    rng = np.random.default_rng(5)
    W = WORKLOADS["D1"]
    n = 160
    p = W.make_params(0, n)
    off = np.arange(n + 1, dtype=np.int64)
    pxy = np.ascontiguousarray(p.T)
    rel = 10.0 ** rng.uniform(-7, -2, n)
    me = rng.choice([0, 49, 50, 1250, 1300, 5000, 20000, 60000, 100000], n).astype(np.uint64)
    ref = O.integrate2_batch(ALGO["QAGP"], W.functor, W.yrange, W.a, W.b, fparams=p, yparams=W.yparams, points=pxy,
                             pts_offsets=off, tol=("rel", 1e-6), max_evals=100000, n=n, rel_tol_each=rel,
                             max_evals_each=me)
    assert len(np.unique(ref["status"])) >= 2
    g = eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=p, yparams=W.yparams, algo=ALGO["QAGP"],
                             tol=Tolerance.Relative(1e-6), max_evals=100000, pts_offsets=off, pts_xy=pxy, n=n,
                             rel_tol_each=rel, max_evals_each=me)
    compare(g, ref, Tolerance.Relative(1e-6), exact=True, label="D1 host")
    g = eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=torch.as_tensor(p, device="cuda:0"),
                             yparams=W.yparams, algo=ALGO["QAGP"], tol=Tolerance.Relative(1e-6), max_evals=100000,
                             pts_offsets=torch.as_tensor(off, device="cuda:0"),
                             pts_xy=torch.as_tensor(pxy, device="cuda:0"), n=n,
                             rel_tol_each=rel, max_evals_each=me.astype(np.int64))
    compare(g, ref, Tolerance.Relative(1e-6), exact=True, label="D1 device")
    # QAGS2 over a rectangle with absolute tolerances
    n = 64
    ab = 10.0 ** rng.uniform(-13, -4, n)
    me = rng.choice([100, 288, 289, 700, 3000, 100000], n).astype(np.uint64)
    ref = O.integrate2_batch(ALGO["QAGS"], "g4", "const", 0.0, 2.0, yparams=[0.0, 3.0], tol=("abs", 1e-9),
                             max_evals=100000, n=n, abs_tol_each=ab, max_evals_each=me)
    g = eng.integrate2_batch("g4", "const", 0.0, 2.0, yparams=[0.0, 3.0], algo=ALGO["QAGS"],
                             tol=Tolerance.Absolute(1e-9), max_evals=100000, n=n, abs_tol_each=ab, max_evals_each=me)
    compare(g, ref, Tolerance.Absolute(1e-9), exact=True, label="g4 host")


def test_qag2_batch_bit_exact(eng):
    """QAG2 (double/algorithm/qag.rs:46-100, deprecated): nests with per-integral budgets, finite and
    infinite rectangles and a dynamic y-range -- bit-identical to the oracle."""
    from gkquad_b200 import Tolerance
    from gpu_common import compare
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    n = 96
    me = rng.choice([10, 288, 289, 700, 3000, 20000, 100000], n).astype(np.uint64)
    cases = [("g4", "const", 0.0, 2.0, [0.0, 3.0], ("abs", 1e-10)),
             ("g4", "const", 0.0, math.inf, [0.0, math.inf], ("abs", 1e-8)),
             ("rsqrt3", "const", -1.0, 1.5, [-2.0, 1.0], ("rel", 1e-9)),
             ("g3", "circle", -0.5, 0.5, [0.25], ("abs", 1e-9))]
    for f, yr, x0, x1, yq, tol in cases:
        ref = O.integrate2_batch(ALGO["QAG"], f, yr, x0, x1, yparams=yq, tol=tol, max_evals=100000, n=n,
                                 max_evals_each=me)
        g = eng.integrate2_batch(f, yr, x0, x1, yparams=yq, algo=ALGO["QAG"], tol=_tol(tol), max_evals=100000, n=n,
                                 max_evals_each=me)
        compare(g, ref, _tol(tol), exact=True, label=f"QAG2 {f}")
    assert len(np.unique(ref["status"])) >= 2


def test_reference_examples(eng):
    """examples/sphere.rs and examples/wallis_integral.rs restated on the host mirror (examples/*.py)."""
    import importlib.util
    from pathlib import Path
    ROOT = Path(__file__).resolve().parent.parent
    mods = {}
    for name in ("sphere", "wallis_integral"):
        spec = importlib.util.spec_from_file_location(name, ROOT / "examples" / f"{name}.py")
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    v1, v2, v3 = mods["sphere"].main()
    assert v1 == 2.0 and abs(v2 - math.pi) <= 1.49e-8 * math.pi * 2 and abs(v3 - 4.0 * math.pi / 3.0) <= 1e-6
    assert abs(mods["wallis_integral"].main(2000) - 36722.362174233774391) < 1.0


def test_api_mirror_on_gpu(eng):
    """The host-side mirror of gkquad's API (single::integral, Integrator builder, integral2)."""
    from gkquad_b200 import Tolerance, double, single
    r = single.integral(single.DeviceIntegrand("square"), (0.0, 1.0)).unwrap()
    assert abs(r.estimate - 1.0 / 3.0) <= 1.49e-8 and r.nevals == 17
    it = single.Integrator.new(single.DeviceIntegrand("recip_p", [1.0])).points([0.0]).tolerance(Tolerance.Absolute(1.49e-8))
    r = it.run((-1.0, 1.0)).unwrap()
    assert abs(r.estimate) <= 1.49e-8 and r.nevals == 250
    r = single.Integrator.new(single.DeviceIntegrand("lorentz_p", [1.0])).run(slice(None, None)).unwrap()
    assert abs(r.estimate - math.pi) <= 1.49e-8
    r = double.integral2(double.DeviceIntegrand2("xy"), ((0.0, 1.0), (0.0, 1.0))).unwrap()
    assert abs(r.estimate - 0.25) <= 1.49e-8 and r.nevals == 289
    q = single.qk17(single.DeviceIntegrand("square"), (0.0, 1.0))
    assert abs(q.estimate - 1.0 / 3.0) < 1e-15
    res = single.Integrator.new(single.DeviceIntegrand("f6")).points([1.0]).tolerance(Tolerance.Absolute(1e-10)).run((0.0, 2.0))
    assert res.has_err() and res.err().name == "Divergent" and res.unwrap_unchecked().nevals == 550
    with pytest.raises(ValueError):
        res.unwrap()


def test_cpp_mirror_reference_benches():
    """The C++ host mirror (gkquad-rs_b200/cpp/gkquad_b200.hpp) reproduces the six reference bench
    assertions (benches/single.rs, benches/double.rs) through gkq_integrate*_batch_host."""
    import subprocess
    exe = Path(__file__).resolve().parent.parent / "gkquad-rs_b200" / "cpp" / "example_benches"
    assert exe.exists(), "build it: make -C gkquad-rs_b200"
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "all reference bench assertions hold" in out.stdout


def test_cpp_pipeline_example():
    """Throughput mode from compiled host code (cpp/example_pipeline.cpp: three handles on three
    streams over the C ABI, device-resident S1 batches): every converged integral of every batch is
    within the requested tolerance of the closed form."""
    import json
    import subprocess
    exe = Path(__file__).resolve().parent.parent / "gkquad-rs_b200" / "cpp" / "example_pipeline"
    assert exe.exists(), "build it: make -C gkquad-rs_b200"
    out = subprocess.run([str(exe), "3", "200000", "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["outside_tolerance_of_closed_form"] == 0 and line["converged_fraction"] > 0.999, line


def test_deferred_join_pipelining(eng):
    """gkq_set_deferred_join: back-to-back independent batches (remainders overlapping the next
    batch's bulk, alternating scratch contexts) give exactly the results of ordered calls."""
    import torch
    from gkquad_b200.workloads import WORKLOADS
    W = WORKLOADS["S3"]  # long tails: many integrals need 3..14 bisection steps
    n = 200_000
    ps = [torch.as_tensor(W.make_params(k * n, n), device="cuda:0") for k in range(5)]
    ref = [eng.integrate_batch(W.functor, 0.0, 1.0, p, algo=ALGO["QAGS"], tol=W.tol, n=n).numpy() for p in ps]
    eng.set_deferred_join(True)
    try:
        outs = [eng.integrate_batch(W.functor, 0.0, 1.0, p, algo=ALGO["QAGS"], tol=W.tol, n=n) for p in ps]
        eng.join()
        torch.cuda.synchronize()
        got = [o.numpy() for o in outs]
    finally:
        eng.set_deferred_join(False)
    for g, r in zip(got, ref):
        assert np.array_equal(g.estimate, r.estimate) and np.array_equal(g.delta, r.delta)
        assert np.array_equal(g.nevals, r.nevals) and np.array_equal(g.status, r.status)


def test_batch_pipeline_matches_single_engine(eng):
    """BatchPipeline (several handles on several streams, deferred join) returns bit for bit what one
    Engine returns for the same batches, whatever the interleaving."""
    import torch
    import gkquad_b200 as gk
    from gkquad_b200.workloads import WORKLOADS
    W = WORKLOADS["S3"]  # heavy tail: plenty of straggler records
    n = 60_000
    nb = 6
    ps = [torch.as_tensor(W.make_params(k * n, n), device="cuda:0") for k in range(nb)]
    want = [eng.integrate_batch(W.functor, W.a, W.b, p, algo=ALGO[W.algo], tol=W.tol, n=n).numpy() for p in ps]
    for depth in (3, 1):
        pipe = gk.BatchPipeline(0, depth=depth)
        try:
            plans = [pipe.plan_batch(k % depth, W.functor, W.a, W.b, ps[k], algo=ALGO[W.algo], tol=W.tol, n=n)
                     for k in range(nb)]
            pipe.fork()
            # 18 batches per lane: two record buffers fill up and are finished on the lane's tail stream
            # (the second one with a grid sized by the first one's record count) before the final join
            for rep in range(3 * depth):
                for k in range(nb):
                    pipe.run(plans[k], k % depth)
            pipe.join()
            torch.cuda.synchronize()
            for k in range(nb):
                g = plans[k].result.numpy()
                assert np.array_equal(g.estimate, want[k].estimate) and np.array_equal(g.delta, want[k].delta)
                assert np.array_equal(g.nevals, want[k].nevals) and np.array_equal(g.status, want[k].status)
            # fresh outputs, one more round: what the tail-stream launches wrote is what is read here
            for k in range(nb):
                r = plans[k].result
                for t in (r.estimate, r.delta, r.nevals, r.status):
                    t.zero_()
            for k in range(nb):
                pipe.run(plans[k], k % depth)
            pipe.join()
            torch.cuda.synchronize()
            for k in range(nb):
                g = plans[k].result.numpy()
                assert np.array_equal(g.estimate, want[k].estimate) and np.array_equal(g.status, want[k].status)
        finally:
            pipe.close()


def test_deferred_remainders_overlap_small_batches(eng):
    """Tiny batches in deferred mode: record buffers fill up and are finished on the tail stream faster
    than a remainder launch runs (75 us), so flushes overlap their predecessors and the engine's
    cudaEventQuery sees "not ready".  Results stay bit-identical to ordered calls."""
    import torch
    import gkquad_b200 as gk
    from gkquad_b200.workloads import WORKLOADS
    W = WORKLOADS["S3"]
    n = 2048
    p = torch.as_tensor(W.make_params(0, n), device="cuda:0")
    want = eng.integrate_batch(W.functor, W.a, W.b, p, algo=ALGO[W.algo], tol=W.tol, n=n).numpy()
    pipe = gk.BatchPipeline(0, depth=1)
    try:
        plan = pipe.plan_batch(0, W.functor, W.a, W.b, p, algo=ALGO[W.algo], tol=W.tol, n=n)
        pipe.fork()
        for _ in range(400):
            pipe.run(plan, 0)
        pipe.join()
        torch.cuda.synchronize()
        g = plan.result.numpy()
        assert np.array_equal(g.estimate, want.estimate) and np.array_equal(g.delta, want.delta)
        assert np.array_equal(g.nevals, want.nevals) and np.array_equal(g.status, want.status)
    finally:
        pipe.close()
