"""Accuracy of the engine's own FP64 exp / sin / cos (gkquad-rs_b200/csrc/gk_math.cuh) against
glibc through numpy, in units in the last place.  The kernels replace libdevice inside the
integrand functors because the rule kernels are issue-bound (profiles/S1_r01.md); they must stay
as accurate as what they replace (CUDA documents 1 ulp for exp and 2 ulp for sin / cos)."""
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
