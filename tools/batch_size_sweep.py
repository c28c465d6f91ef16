"""Throughput of the S1 hot path as a function of the batch size (device-resident, deferred join).

    python tools/batch_size_sweep.py
"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]
import gkquad_b200 as gk  # noqa: E402
from gkquad_b200.workloads import WORKLOADS  # noqa: E402

W = WORKLOADS["S1"]
eng = gk.Engine(0)
print("| batch | ms / batch (device) | integrals/s | ordered-mode ms | host-entry ms |")
print("|---|---|---|---|---|")
for n in (1 << 10, 1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24):
    p = W.make_params(0, n)
    pd = torch.as_tensor(p, device="cuda:0")
    plans = [eng.plan_batch(W.functor, W.a, W.b, pd, algo=1, tol=W.tol, n=n) for _ in range(4)]
    reps = max(8, min(512, (1 << 26) // n))

    def timed(deferred):
        eng.set_deferred_join(deferred)
        for k in range(8):
            plans[k % 4].run()
        eng.join()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(reps):
            plans[k % 4].run()
        eng.join()
        e1.record()
        torch.cuda.synchronize()
        eng.set_deferred_join(False)
        return e0.elapsed_time(e1) / reps

    ms_d, ms_o = timed(True), timed(False)
    hp = eng.plan_batch(W.functor, W.a, W.b, p, algo=1, tol=W.tol, n=n)
    for _ in range(3):
        hp.run()
    t0 = time.perf_counter()
    hreps = max(3, min(50, (1 << 24) // n))
    for _ in range(hreps):
        hp.run()
    ms_h = (time.perf_counter() - t0) / hreps * 1e3
    print(f"| {n} | {ms_d:.4f} | {n / ms_d * 1e3:.3e} | {ms_o:.4f} | {ms_h:.4f} |")
