"""Device-entry time of workload families on cuda:0 (no oracle): ms per call, evaluations/s and a
checksum of the results, so that tuning builds (GKQ_LIB=...) can be compared quickly.

    python tools/quick_time.py P0 P1 I1 D0:100000 D1:16384
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]
import gkquad_b200 as gk  # noqa: E402
from gkquad_b200.workloads import WORKLOADS  # noqa: E402

ALGO = {"AUTO": 0, "QAGS": 1, "QAGP": 2}
eng = gk.Engine(0)
for spec in sys.argv[1:] or ["S1", "P0", "P1", "I1"]:
    name, _, nn = spec.partition(":")
    W = WORKLOADS[name]
    n = int(nn) if nn else (1 << 20 if W.dim == 1 else 100_000)
    p = W.make_params(0, n)
    pd = torch.as_tensor(p, device="cuda:0")
    kw = dict(algo=ALGO[W.algo], tol=W.tol, max_evals=W.max_evals, n=n)
    if W.dim == 1:
        if W.points == "param0":
            kw.update(pts_offsets=torch.arange(n + 1, dtype=torch.int64, device="cuda:0"), pts=pd[0].contiguous())
        elif W.points:
            kw["points"] = W.points
        plan = eng.plan_batch(W.functor, W.a, W.b, pd, **kw)
        run = plan.run
    else:
        if W.points == "param01":
            kw.update(pts_offsets=torch.arange(n + 1, dtype=torch.int64, device="cuda:0"), pts_xy=pd.t().contiguous())
        elif W.points:
            kw["points"] = W.points

        def run():
            return eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=pd, yparams=W.yparams, **kw)
    for _ in range(2):
        r = run()
    torch.cuda.synchronize()
    reps = 5 if W.dim == 1 else 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    g = r.numpy()
    evals = float(g.nevals.astype(np.float64).sum())
    h = hashlib.sha1(g.estimate.tobytes() + g.delta.tobytes() + g.nevals.tobytes() + g.status.tobytes()).hexdigest()[:12]
    print(f"{name:3s} n {n:8d}  {ms:9.4f} ms  {n / ms / 1e3:9.3e} integrals/s  {evals / ms / 1e3:9.3e} evals/s  "
          f"mean nevals {evals / n:9.2f}  status {np.bincount(g.status, minlength=6).tolist()}  sha {h}", flush=True)
