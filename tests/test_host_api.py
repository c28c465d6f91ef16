"""Host-side mirror of gkquad's public types (no GPU needed): Tolerance, Range, builders,
argument validation -- the behaviour the reference documents in common.rs, single/common.rs,
single/integrator.rs, double/range.rs."""
import math

import numpy as np
import pytest

from gkquad_b200 import IntegrationResult, RuntimeError_, Solution, Tolerance, double, single
from gkquad_b200.workloads import WORKLOADS, block_cyclic_shard, splitmix64, uniform


def test_tolerance_to_abs_and_default():
    assert Tolerance.default() == Tolerance.AbsOrRel(1.49e-8, 1.49e-8)  # common.rs:44-49
    assert Tolerance.Absolute(1e-3).to_abs(5.0) == 1e-3
    assert Tolerance.Relative(1e-3).to_abs(5.0) == 5e-3
    assert Tolerance.AbsOrRel(1e-2, 1e-3).to_abs(5.0) == 1e-2
    assert Tolerance.AbsAndRel(1e-2, 1e-3).to_abs(5.0) == 5e-3
    assert Tolerance.Relative(float("nan")).contains_nan()
    assert not Tolerance.Absolute(1.0).contains_nan()
    assert Tolerance.AbsOrRel(1.0, float("nan")).contains_nan()


def test_range_rules():
    with pytest.raises(ValueError):
        single.Range(float("nan"), 1.0)  # single/common.rs:94
    assert single.Range.new(float("nan"), 1.0) is None
    r = single.Range.from_(slice(None, None))  # `..`
    assert (r.begin, r.end) == (-math.inf, math.inf)
    r = single.Range.from_(slice(0.0, None))  # `0.0..`
    assert (r.begin, r.end) == (0.0, math.inf)
    assert single.Range.from_((1.0, 0.0)).begin == 1.0  # reversed ranges are legal
    with pytest.raises(ValueError):
        double.Rectangle.new(0.0, math.inf, 0.0, 1.0)  # double/range.rs:20-22
    assert double.Rectangle.new(0.0, float("nan"), 0.0, 1.0) is None
    rect = double.Rectangle.from_(((0.0, None), (0.0, None)))  # From<(R1, R2)> allows infinite sides
    assert rect.xrange.end == math.inf and rect.yrange.end == math.inf


def test_builder_validation_and_defaults():
    f = single.DeviceIntegrand("square")
    it = single.Integrator.new(f)
    assert isinstance(it.get_algorithm(), single.AUTO)
    assert it.config.max_evals == 2000 and it.config.points == [] and it.config.tolerance == Tolerance.default()
    with pytest.raises(AssertionError, match="Tolerance must not contain NAN"):
        it.tolerance(Tolerance.Relative(float("nan")))  # integrator.rs:71
    with pytest.raises(AssertionError, match="cannot include NAN value in singular points"):
        it.points([0.0, float("nan")])  # integrator.rs:86-90
    it2 = it.max_evals(500).points([0.5]).algorithm(single.QAGP())
    assert it2.config.max_evals == 500 and it2.config.points == [0.5] and it2.get_algorithm() == single.QAGP()
    assert single.QAGS() == single.QAGS() and single.QAGS() != single.QAGP()  # extra_traits!
    i2 = double.Integrator2.new(double.DeviceIntegrand2("xy"))
    assert i2.config.max_evals == 100000 and isinstance(i2.get_algorithm(), double.AUTO2)
    with pytest.raises(AssertionError):
        i2.points([(0.0, float("nan"))])


def test_value_with_error_semantics():
    ok = IntegrationResult(Solution(1.0, 1e-9, 17))
    bad = IntegrationResult(Solution(2.0, 3.0, 1967), RuntimeError_.Divergent)
    assert ok.unwrap().estimate == 1.0 and not ok.has_err() and ok.err() is None
    assert bad.has_err() and bad.err() == RuntimeError_.Divergent and bad.ok() is None
    assert bad.unwrap_unchecked().estimate == 2.0  # the partial value is always there
    with pytest.raises(ValueError):
        bad.unwrap()
    assert Solution().delta == 1.7976931348623157e308 and Solution().nevals == 0  # common.rs:222-231
    assert [e.value for e in RuntimeError_] == [1, 2, 3, 4, 5]  # error.rs:145-176 declaration order


def test_parameter_generator_is_counter_based():
    # splitmix64 reference values (public test vector for seed 0: first outputs of the stream)
    z = splitmix64(np.array([0], dtype=np.uint64))
    assert int(z[0]) == 0xE220A8397B1DCDAF
    u = uniform(0, 0, 1000)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert np.array_equal(uniform(1, 10, 5), uniform(1, 0, 15)[10:])  # position-independent
    p = WORKLOADS["S1"].make_params(0, 4096)
    assert p.shape == (2, 4096) and p[0].min() >= 0.5 and p[0].max() <= 2.0 and p[1].max() <= 10.0


def test_block_cyclic_shard_partitions_the_batch():
    n = 3 * 4096 * 4 + 123
    seen = np.concatenate([block_cyclic_shard(n, r, 4) for r in range(4)])
    assert np.array_equal(np.sort(seen), np.arange(n))
    assert block_cyclic_shard(n, 1, 4)[0] == 4096
