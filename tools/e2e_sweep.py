"""End-to-end (pinned host buffers) throughput of gkq_integrate_batch_host on S1, 1 M integrals per call:
host threads (one handle each) x chunk schedule.

    python tools/e2e_sweep.py
"""
import os
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
NT = 8
engs = [gk.Engine(0) for _ in range(NT)]
plans, outs = [], []
for t in range(NT):
    ph = torch.as_tensor(W.make_params(t * n, n)).pin_memory()
    o = (torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
         torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory())
    out = (o[0].numpy(), o[1].numpy(), o[2].numpy().view(np.uint32), o[3].numpy().view(np.uint32))
    outs.append((ph, o, out))
    plans.append(engs[t].plan_batch(W.functor, W.a, W.b, ph.numpy(), algo=1, tol=W.tol, max_evals=W.max_evals, n=n, out=out))


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
    return dt / (nthreads * calls_per_thread) * 1e3


for ch, c0 in ((0, 0), (262144, 262144), (524288, 524288), (1048576, 1048576)):
    if ch:
        os.environ["GKQ_HOST_CHUNK"] = str(ch)
        os.environ["GKQ_HOST_CHUNK0"] = str(c0)
    for p in plans:
        p.run()
    row = []
    for nt in (1, 2, 3, 4, 6, 8):
        bench(nt, 5)
        ms = min(bench(nt, 40) for _ in range(2))
        row.append(f"{nt} thr {ms:.3f} ms ({n / ms * 1e-6:.3f}e9/s)")
    print(("schedule by calls in flight (default)" if not ch else f"chunk {ch:8d} first {c0:8d}") + ": " + "  ".join(row), flush=True)
