"""Accuracy of the engine's own FP64 exp / sin / cos (gkquad-rs_b200/csrc/gk_math.cuh) against
glibc through numpy, in units in the last place.  The kernels replace libdevice inside the
integrand functors because the rule kernels are issue-bound (profiles/S1_r01.md); they must stay
as accurate as what they replace (CUDA documents 1 ulp for exp and 2 ulp for sin / cos)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ulp_err(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    ulp = np.spacing(np.abs(ref))
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(got - ref) / ulp
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
    return np.where(same, 0.0, e)


@pytest.fixture(scope="module")
def eng():
    import gkquad_b200 as gk
    return gk.default_engine(0)


def _eval(eng, functor, x, p0=1.0):
    import torch
    y = eng.eval_batch(functor, torch.as_tensor(x, device="cuda:0"),
                       torch.as_tensor(np.full((1, len(x)), p0), device="cuda:0"))
    return y.cpu().numpy()


def _grid(lo, hi, n, seed):
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(lo, hi, n), np.linspace(lo, hi, 1001)])


def test_exp_accuracy(eng):
    x = np.concatenate([_grid(-700, 700, 400_000, 1), _grid(-2, 2, 400_000, 2), _grid(-1e-3, 1e-3, 10_000, 3),
                        [0.0, -0.0, 707.9, -707.9, 708.5, 709.7, 710.0, -745.0, -746.0, -720.0, np.inf, -np.inf, np.nan,
                         0.5 * np.log(2), -0.5 * np.log(2), 1e-320]])
    got = _eval(eng, "exp_p", x)
    ref = np.exp(x)
    e = _ulp_err(got, ref)
    print("exp: max ulp error", e.max(), "mean", e.mean())
    assert e.max() <= 1.0
    assert np.isnan(got[np.isnan(x)]).all() and got[x == -np.inf][0] == 0.0 and got[x == np.inf][0] == np.inf


@pytest.mark.parametrize("fn,ref", [("cos_p", np.cos), ("sin_p", np.sin)])
def test_sincos_accuracy(eng, fn, ref):
    k = np.arange(-2000, 2001)
    near_roots = np.concatenate([k * (np.pi / 2), k * (np.pi / 2) + 1e-9, np.nextafter(k * (np.pi / 2), np.inf)])
    x = np.concatenate([_grid(-40, 40, 400_000, 4), _grid(-1e6, 1e6, 400_000, 5), _grid(-1e-4, 1e-4, 10_000, 6),
                        near_roots, [0.0, 1e-300, 1.0737e9, 1.1e9, 1e15, -1e22, np.inf, np.nan]])
    got = _eval(eng, fn, x)
    with np.errstate(invalid="ignore"):
        want = ref(x)
    e = _ulp_err(got, want)
    print(fn, "max ulp error", e.max(), "mean", e.mean(), "argmax", x[np.argmax(e)])
    assert e.max() <= 2.0
    assert np.isnan(got[-1]) and np.isnan(got[-2])


def _libm(fn, x):
    """glibc's scalar libm through Python's math module (numpy's own log / power use a vectorised
    libm with ~1 ulp error on AVX-512 hosts, which is not the oracle's arithmetic)."""
    out = np.empty(len(x))
    for i, v in enumerate(x.tolist()):
        try:
            out[i] = fn(v)
        except (ValueError, ZeroDivisionError):  # log(0), 0 ** -p, log(-x)
            out[i] = -np.inf if fn is math.log and v == 0.0 else (np.inf if v == 0.0 else np.nan)
        except OverflowError:
            out[i] = np.inf
    return out


def _eval2(eng, functor, x, params):
    import torch
    p = np.ascontiguousarray(np.broadcast_to(np.asarray(params, dtype=np.float64)[:, None], (len(params), len(x))))
    return eng.eval_batch(functor, torch.as_tensor(x, device="cuda:0"), torch.as_tensor(p, device="cuda:0")).cpu().numpy()


def test_log_accuracy(eng):
    """gk_log_fast (table-driven, csrc/gk_math_logpow.cuh) through the functor ln|x - c| with c = 0:
    <= 0.53 ulp on the host check (tools/hostcheck/check_logpow.cpp); here against glibc (<= 1 ulp allows
    for glibc's own 0.52).  Non-normal / non-positive arguments take the libdevice path."""
    rng = np.random.default_rng(11)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 400_000)), rng.uniform(0.5, 1.5, 400_000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 100_000), rng.uniform(0, 1, 200_000), -rng.uniform(0, 5, 1000),
                        [1.0, 2.0, 0.5, 1e-310, 5e-324, 0.0, np.inf, np.nan, np.nextafter(1.0, 0), np.nextafter(1.0, 2)]])
    got = _eval2(eng, "logabs_p", x, [0.0])
    want = _libm(math.log, np.abs(x))  # glibc through the math module (numpy may use a vector libm)
    e = _ulp_err(got, want)
    fin = np.isfinite(want)
    print("log: max ulp error", e[fin].max(), "mean", e[fin].mean(), "bit-equal share", float((got == want)[fin].mean()))
    assert e[fin].max() <= 1.0
    assert got[x == 0.0][0] == -np.inf and got[np.isinf(x)][0] == np.inf and np.isnan(got[np.isnan(x)]).all()
    assert (got == want)[fin].mean() > 0.997


@pytest.mark.parametrize("p", [0.1, 0.35, 0.6, 0.9, -2.6, 3.0, 17.25])
def test_pow_accuracy(eng, p):
    """gk_pow_fast through the functor |x - c|^-p with c = 0 (P1: p in [0.1, 0.6]; f1 / f2: x^2.6, x^-0.9)."""
    rng = np.random.default_rng(int(p * 100) + 1000)
    x = np.concatenate([rng.uniform(0, 1, 400_000), np.exp(rng.uniform(-30, 30, 300_000)), 1.0 + rng.uniform(-1e-4, 1e-4, 50_000),
                        [1.0, 2.0, 0.25, 1e-300, 1e300, 1e-320, 0.0, np.inf, np.nan]])
    got = _eval2(eng, "abspow_p", x, [0.0, p])
    want = _libm(lambda v: math.pow(v, -p), np.abs(x))
    e = _ulp_err(got, want)
    fin = np.isfinite(want) & (want != 0) & (np.abs(want) > 1e-300)
    print("pow p =", p, "max ulp error", e[fin].max(), "mean", e[fin].mean(), "bit-equal share", float((got == want)[fin].mean()))
    assert e[fin].max() <= 1.0
    assert np.isnan(got[np.isnan(x)]).all()
    i0 = np.nonzero(x == 0.0)[0][0]
    assert got[i0] == want[i0]
    assert (got == want)[fin].mean() > 0.995


def _edge_values():
    """Operands around every threshold of the fast paths: zeros, denormals, the ends of the exponent
    range, infinities, NaN, and values whose quotients / roots land in the denormal range."""
    tiny = np.array([0.0, -0.0, 5e-324, 1e-310, 2.2250738585072014e-308, 4.4501477170144023e-308, 1e-300, 1e-292, 1e-290,
                     6e-293, 7e-293, 1e-200, 1e-154, 1e-153, 1e-100])
    mid = np.array([1e-10, 0.1, 1.0 / 3.0, 0.5, 1.0, 1.5, 2.0, 3.0, 10.0, 1e10])
    huge = np.array([1e100, 1e153, 1e154, 1e200, 1e290, 1e300, 8.98846567431158e307, 1.7976931348623157e308, np.inf, np.nan])
    v = np.concatenate([tiny, mid, huge])
    return np.concatenate([v, -v])


def test_branch_free_div_sqrt(eng):
    """gk_div_fast / gk_sqrt_fast (csrc/gk_math.cuh) are nvcc's own fast paths of `/` and `sqrt` with the
    slow-path branch turned into a flag: the sample path of the rule kernels (gkq_functor_eval_batch_ex,
    mode 1) must return exactly IEEE's correctly rounded quotient / root -- what numpy computes on the host --
    for random operands over the whole exponent range AND for every edge operand (where the flag sends the
    sample through the compiler's own `/` and `sqrt`)."""
    rng = np.random.default_rng(7)
    n = 1 << 21
    with np.errstate(all="ignore"):
        # p / x : random magnitudes over the whole exponent range, then the cross product of the edge values
        num = np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(-1060, 1020, n)) * rng.choice([-1.0, 1.0], n)
        den = np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(-1060, 1020, n)) * rng.choice([-1.0, 1.0], n)
        ev = _edge_values()
        en, ed = np.meshgrid(ev, ev)
        num = np.concatenate([num, en.ravel(), rng.uniform(0.5, 2.0, n)])
        den = np.concatenate([den, ed.ravel(), rng.uniform(-1.0, 1.0, n)])
        want = num / den
        for mode in (0, 1):
            got = eng.eval_batch("recip_p", den, params=num[None, :], mode=mode).cpu().numpy()
            same = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
            assert same.all(), (mode, int((~same).sum()), num[~same][:4], den[~same][:4], got[~same][:4], want[~same][:4])
        # sqrt(x - p): radicands over the whole range, negative ones, zeros, denormals, inf, nan
        x = np.concatenate([np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(-1074, 1023, n)), np.abs(ev), -np.abs(ev),
                            rng.uniform(-1.0, 4.0, n)])
        want = np.sqrt(x)
        for mode in (0, 1):
            got = eng.eval_batch("sqrt_shift_p", x, params=np.zeros((1, x.size)), mode=mode).cpu().numpy()
            same = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
            assert same.all(), (mode, int((~same).sum()), x[~same][:4], got[~same][:4], want[~same][:4])
        # a body with both: 1 / (1 + p x^2) and the 2-D bodies run through the full-size parity tests


def test_sample_path_equals_always_valid_path(eng):
    """mode 1 (branch-free first) == mode 0 (always valid) bit for bit for every registered 1-D functor, on
    arguments inside and far outside the fast ranges."""
    rng = np.random.default_rng(11)
    n = 1 << 18
    x = np.concatenate([rng.uniform(-3.0, 3.0, n), rng.uniform(-1e6, 1e6, n // 4), np.ldexp(rng.uniform(1, 2, n // 4), rng.integers(-1070, 1020, n // 4)),
                        _edge_values()])
    for fid in range(eng.L.gkq_functor_count(1)):
        _, name, npar, _ = eng.functor(1, fid)
        p = rng.uniform(0.25, 3.0, (max(npar, 1), x.size))
        if name == "sinpow_p":
            p[0] = rng.integers(-6, 7, x.size)
        a = eng.eval_batch(name, x, params=p if npar else None, mode=0).cpu().numpy()
        b = eng.eval_batch(name, x, params=p if npar else None, mode=1).cpu().numpy()
        same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), (name, int((~same).sum()), x[~same][:4], a[~same][:4], b[~same][:4])
