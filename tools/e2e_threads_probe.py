"""End-to-end (host buffers) throughput of gkq_integrate_batch_host with 1, 2 and 3 host threads, each
driving its own handle (the call is blocking and releases the GIL): H2D of one call overlaps D2H of
another."""
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

W = WORKLOADS["S1"]
n = 1 << 20
NT = 3
engs = [gk.Engine(0) for _ in range(NT)]
plans = []
outs = []
for t in range(NT):
    ph = torch.as_tensor(W.make_params(t * n, n)).pin_memory()
    o = (torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
         torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory())
    out = (o[0].numpy(), o[1].numpy(), o[2].numpy().view(np.uint32), o[3].numpy().view(np.uint32))
    outs.append((ph, o, out))
    plans.append(engs[t].plan_batch(W.functor, W.a, W.b, ph.numpy(), algo=1, tol=W.tol, max_evals=W.max_evals, n=n, out=out))
for p in plans:
    for _ in range(3):
        p.run()


def bench(nthreads, calls_per_thread):
    def work(t):
        for _ in range(calls_per_thread):
            plans[t].run()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    total = nthreads * calls_per_thread
    conv = float((outs[0][2][3] == 0).sum())
    print(f"threads={nthreads}: {dt / total * 1e3:.3f} ms per 1M batch, {conv * total / dt:.4g} converged integrals/s")


for nt in (1, 2, 3, 1, 2, 3):
    bench(nt, 60)
