"""examples/wallis_integral.rs of the reference, through the B200 engine.

The reference runs ONE Integrator 10 times per sweep (m = 1..10) over sin(pi x)^m on [0, 1] with
Tolerance::Absolute(1.) and repeats the sweep 100 000 times to time it.  Here a sweep is one batch:
the 10 integrals (or 10 x `sweeps` of them) go to the GPU in a single call.
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "gkquad-rs_b200")]
from gkquad_b200 import Tolerance, single  # noqa: E402

EXPECTED = 36722.362174233774391


def main(sweeps: int = 100000):
    m = np.tile(np.arange(1, 11, dtype=np.float64), sweeps)[None, :]  # [1, 10 * sweeps]
    f = single.DeviceIntegrand("sinpow_p", m)
    it = single.Integrator.new(f).tolerance(Tolerance.Absolute(1.0))
    it.run((0.0, 1.0))  # warm-up (allocations)
    t0 = time.perf_counter()
    res = it.run((0.0, 1.0))
    dt = time.perf_counter() - t0
    est = np.asarray(res.estimate).reshape(sweeps, 10)
    inv = 1.0 / np.prod(est, axis=1)
    assert np.all(np.abs(inv - EXPECTED) < 1.0), inv[:3]
    print(f"Elapsed time per sweep of 10 integrals: {dt / sweeps * 1e9:.1f} ns ({sweeps} sweeps in one batch)")
    return float(inv[0])


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 100000)
