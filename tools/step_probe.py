"""Per-step GPU time and host enqueue time of the device entry, ordered vs deferred join."""
import sys
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
eng = gk.Engine(0)
NS = 6
ps = [torch.as_tensor(W.make_params(k * n, n), device="cuda:0") for k in range(NS)]
outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda"),
         torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")) for _ in range(NS)]


plans = [eng.plan_batch(W.functor, 0.0, 1.0, ps[k], algo=1, tol=W.tol, n=n, out=outs[k]) for k in range(NS)]


def run(K, deferred):
    eng.set_deferred_join(deferred)
    for k in range(12):
        plans[k % NS].run()
    eng.join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for k in range(K):
        plans[k % NS].run()
    eng.join()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    eng.set_deferred_join(False)
    print(f"deferred={deferred}: GPU {e0.elapsed_time(e1) / K * 1e3:.1f} us/step, host enqueue {(t1 - t0) / K * 1e6:.1f} us/step, "
          f"launches {eng.stats()['kernel_launches']}")


for d in (False, True, False, True):
    run(40, d)
