"""gkquad_b200 -- host-side mirror of the gkquad crate's public surface for the batched
B200 engine (reference re-exports: gkquad/src/lib.rs:46-55, prelude.rs:1-7).

    from gkquad_b200 import single, double, Tolerance
    r = single.integral(single.DeviceIntegrand("square"), (0.0, 1.0)).unwrap()

All computation happens in libgkquad_b200.so (hand-written sm_100a kernels behind the C ABI in
include/gkquad_b200.h); there is no CPU fallback.
"""
from . import double, single  # noqa: F401
from .common import (BatchResult, IntegrationResult, RuntimeError, RuntimeError_, Solution,  # noqa: F401
                     Tolerance)
from .engine import ALGO_AUTO, ALGO_QAG, ALGO_QAGP, ALGO_QAGS, BatchPlan, Engine, default_engine  # noqa: F401
from ._ffi import GkqError, functor_table  # noqa: F401
from .pipeline import BatchPipeline  # noqa: F401
