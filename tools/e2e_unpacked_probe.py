"""Host entry, S1, 1 M integrals per call: nevals / status packed into one word on the device and unpacked by the
calling thread (default) against four separate result copies (GKQ_HOST_UNPACKED=1), by host threads."""
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
NT = 6
engs = [gk.Engine(0) for _ in range(NT)]
plans = []
for t in range(NT):
    ph = torch.as_tensor(W.make_params(t * n, n)).pin_memory()
    o = (torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
         torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory())
    out = (o[0].numpy(), o[1].numpy(), o[2].numpy().view(np.uint32), o[3].numpy().view(np.uint32))
    plans.append((engs[t].plan_batch(W.functor, W.a, W.b, ph.numpy(), algo=1, tol=W.tol, max_evals=W.max_evals, n=n, out=out), ph, o))


def bench(nthreads, calls):
    def work(t):
        for _ in range(calls):
            plans[t][0].run()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return (time.perf_counter() - t0) / (nthreads * calls) * 1e3


for rep in range(2):
    for mode in ("packed", "unpacked"):
        if mode == "unpacked":
            os.environ["GKQ_HOST_UNPACKED"] = "1"
        else:
            os.environ.pop("GKQ_HOST_UNPACKED", None)
        row = []
        for nt in (1, 2, 4, 6):
            bench(nt, 5)
            ms = min(bench(nt, 40) for _ in range(2))
            row.append(f"{nt} thr {ms:.3f} ms ({n / ms * 1e-6:.3f}e9/s)")
        print(f"{mode:9s}: " + "  ".join(row), flush=True)
