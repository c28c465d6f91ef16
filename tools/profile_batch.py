"""Runs one workload a few times on cuda:0 -- the short command that ncu wraps
(B200_PROFILING.md: run it plain first, then under ncu with the same arguments).

    python tools/profile_batch.py S1 [batch] [reps]
    python tools/profile_batch.py eval:<functor> [n]      # functor-only micro-kernel (flop counts)
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]
import gkquad_b200 as gk  # noqa: E402
from gkquad_b200.workloads import WORKLOADS, csr_points_from_param0  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "S1"
eng = gk.Engine(0)
if what.startswith("eval:"):
    names = what[5:].split(",")
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    for name in names:
        fid, nm, npar, fl = eng.functor(1, name)
        W = next((w for w in WORKLOADS.values() if w.functor == name and w.dim == 1), None)
        rng = np.random.default_rng(0)
        if W is not None:
            p = W.make_params(0, n)
            lo, hi = (W.a, W.b) if np.isfinite(W.a) and np.isfinite(W.b) else (-20.0, 20.0)
        else:
            p = np.ones((max(npar, 1), n))
            lo, hi = 0.01, 2.0
        x = torch.as_tensor(lo + (hi - lo) * rng.random(n), device="cuda:0")
        y = eng.eval_batch(name, x, torch.as_tensor(p, device="cuda:0") if npar else None)
        torch.cuda.synchronize()
        print(name, "n", n, "sum", float(y.sum()))
    sys.exit(0)

W = WORKLOADS[what]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
algo = {"AUTO": 0, "QAGS": 1, "QAGP": 2}[W.algo]
p = W.make_params(0, n)
pd = torch.as_tensor(p, device="cuda:0")
kw = dict(algo=algo, tol=W.tol, max_evals=W.max_evals, n=n)
if W.dim == 1:
    if W.points == "param0":
        off, pts = csr_points_from_param0(p)
        kw.update(pts_offsets=torch.as_tensor(off, device="cuda:0"), pts=torch.as_tensor(pts, device="cuda:0"))
    elif W.points:
        kw["points"] = W.points
    for _ in range(reps):
        r = eng.integrate_batch(W.functor, W.a, W.b, pd, **kw)
else:
    if W.points == "param01":
        kw.update(pts_offsets=torch.arange(n + 1, dtype=torch.int64, device="cuda:0"),
                  pts_xy=torch.as_tensor(np.ascontiguousarray(p.T), device="cuda:0"))
    elif W.points:
        kw["points"] = W.points
    for _ in range(reps):
        r = eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=pd, yparams=W.yparams, **kw)
torch.cuda.synchronize()
g = r.numpy()
print(what, "n", n, "mean nevals", g.nevals.mean(), "status", np.bincount(g.status), "launches",
      eng.stats()["kernel_launches"])
