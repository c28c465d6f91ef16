"""2-D nests of one shard as ONE level-synchronous call or as L concurrent calls (L handles, L streams,
L host threads): do the round tails of one call overlap the other's kernels?

    python tools/two_lane_2d_probe.py D1:12500 D0:12500 D1:100000
"""
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]
import gkquad_b200 as gk  # noqa: E402
from gkquad_b200.workloads import WORKLOADS  # noqa: E402

ALGO = {"AUTO": 0, "QAGS": 1, "QAGP": 2}
dev = "cuda:0"
engines = [gk.Engine(0) for _ in range(4)]
streams = [torch.cuda.Stream(device=dev) for _ in range(4)]


def make_call(e, st, W, p_host):
    m = p_host.shape[1]
    p_dev = torch.as_tensor(np.ascontiguousarray(p_host), device=dev)
    kw = dict(algo=ALGO[W.algo], tol=W.tol, max_evals=W.max_evals, n=m)
    if W.points == "param01":
        kw["pts_offsets"] = torch.arange(m + 1, dtype=torch.int64, device=dev)
        kw["pts_xy"] = p_dev.t().contiguous()
    elif W.points:
        kw["points"] = W.points

    def run():
        with torch.cuda.stream(st):
            return e.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=p_dev, yparams=W.yparams, **kw)
    return run


for spec in sys.argv[1:] or ["D1:12500", "D0:12500"]:
    name, _, nn = spec.partition(":")
    W = WORKLOADS[name]
    n = int(nn)
    p = W.make_params(0, n)
    ref = None
    for L in (1, 2, 3, 4):
        bounds = [n * k // L for k in range(L + 1)]
        calls = [make_call(engines[k], streams[k], W, p[:, bounds[k]:bounds[k + 1]]) for k in range(L)]
        res = [None] * L

        def worker(k):
            res[k] = calls[k]()

        def once():
            th = [threading.Thread(target=worker, args=(k,)) for k in range(1, L)]
            for t_ in th:
                t_.start()
            worker(0)
            for t_ in th:
                t_.join()
            torch.cuda.synchronize()
        once()
        once()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        dt = (time.perf_counter() - t0) / reps
        est = np.concatenate([r.estimate.cpu().numpy() for r in res])
        nev = np.concatenate([r.nevals.cpu().numpy() for r in res]).astype(np.int64)
        if ref is None:
            ref = (est, nev)
        same = bool(np.array_equal(est, ref[0]) and np.array_equal(nev, ref[1]))
        print(f"{name} n {n:7d} lanes {L}: {dt * 1e3:9.3f} ms  {nev.sum() / dt:.4g} evals/s  bit-equal to 1 lane: {same}", flush=True)
