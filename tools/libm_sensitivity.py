"""How often does gkquad's OWN algorithm change its status / nevals when the integrand's libm is
one ulp off?  Runs the pinned CPU oracle twice on workload S1 -- glibc's cos, and glibc's cos moved
one ulp up (oracle-only functor `expcos_p_ulp`) -- and counts the differences.  This is the
noise floor below which "identical status" cannot be asked of any implementation whose exp / cos
are not bit-identical to glibc's (tests/gpu_common.compare).
Usage: python tools/libm_sensitivity.py [n]      (CPU only; uses oracle/)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gkquad-rs_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as O  # noqa: E402
from gkquad_b200.workloads import WORKLOADS  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
for wname in ("S1", "S3"):
    W = WORKLOADS[wname]
    p = W.make_params(0, n)
    kw = dict(params=p, tol=("rel", 1e-10), max_evals=W.max_evals, n=n)
    r0 = O.integrate_batch(O.ALGO_QAGS, "expcos_p", W.a, W.b, **kw)
    r1 = O.integrate_batch(O.ALGO_QAGS, "expcos_p_ulp", W.a, W.b, **kw)
    st = int((r0["status"] != r1["status"]).sum())
    nv = int((r0["nevals"] != r1["nevals"]).sum())
    ok = (r0["status"] == 0) & (r1["status"] == 0)
    worst = float(np.max(np.abs(r0["estimate"] - r1["estimate"])[ok] / (1e-10 * np.abs(r0["estimate"])[ok])))
    print(f"{wname}: n={n}  status flips {st}  nevals flips {nv}  "
          f"max |dI| / (rel_tol |I|) over integrals converged in both = {worst:.3g}")
