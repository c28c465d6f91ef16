"""Throughput + parity of every BASELINE.json workload family on one GPU, next to the CPU oracle.

    python tools/workload_sweep.py [names...]      # default: all
Prints one JSON line per workload and a markdown table (goes into profiles/).
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200"), str(ROOT / "tests")]
import gkquad_b200 as gk  # noqa: E402
from gkquad_b200.workloads import WORKLOADS, csr_points_from_param0  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker + CPU baseline, not product)

ALGO = {"AUTO": 0, "QAGS": 1, "QAGP": 2}
SIZES = {"S1": 1 << 20, "S2": 1 << 20, "S3": 1 << 20, "P0": 1 << 20, "P1": 1 << 20, "P2": 1 << 20,
         "I0": 1 << 20, "I1": 1 << 20, "I2": 1 << 20, "D0": 100_000, "D1": 8192}
CPU_SAMPLE = {"P1": 1 << 18, "P2": 1 << 18, "I2": 1 << 18, "S3": 1 << 19, "D0": 20_000, "D1": 2048}


def tol_to_oracle(t):
    return {0: ("abs", t.abs_tol), 1: ("rel", t.rel_tol), 2: ("abs_or_rel", t.abs_tol, t.rel_tol),
            3: ("abs_and_rel", t.abs_tol, t.rel_tol)}[t.kind]


def main():
    names = sys.argv[1:] or list(SIZES)
    eng = gk.Engine(0)
    peak, _ = eng.measure_fp64_peak(3)
    rows = []
    for name in names:
        W = WORKLOADS[name]
        n = SIZES[name]
        p = W.make_params(0, n)
        pd = torch.as_tensor(p, device="cuda:0")
        kw = dict(algo=ALGO[W.algo], tol=W.tol, max_evals=W.max_evals, n=n)
        okw = dict(tol=tol_to_oracle(W.tol), max_evals=W.max_evals)
        if W.dim == 1:
            if W.points == "param0":
                off, pts = csr_points_from_param0(p)
                kw.update(pts_offsets=torch.as_tensor(off, device="cuda:0"), pts=torch.as_tensor(pts, device="cuda:0"))
            elif W.points:
                kw["points"] = W.points
            plan = eng.plan_batch(W.functor, W.a, W.b, pd, **kw)
            run = plan.run
        else:
            if W.points == "param01":
                kw.update(pts_offsets=torch.arange(n + 1, dtype=torch.int64, device="cuda:0"),
                          pts_xy=torch.as_tensor(np.ascontiguousarray(p.T), device="cuda:0"))
            elif W.points:
                kw["points"] = W.points

            def run():
                return eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=pd, yparams=W.yparams, **kw)
        for _ in range(2):
            r = run()
        torch.cuda.synchronize()
        reps = 5 if W.dim == 1 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        g = r.numpy()
        # CPU oracle on a bounded sample of the same batch
        m = min(n, CPU_SAMPLE.get(name, n))
        t0 = time.perf_counter()
        if W.dim == 1:
            ck = dict(okw)
            if W.points == "param0":
                off, pts = csr_points_from_param0(p[:, :m])
                ck.update(pts_offsets=off, points=pts)
            elif W.points:
                ck["points"] = W.points
            ref = O.integrate_batch(ALGO[W.algo], W.functor, W.a, W.b, params=np.ascontiguousarray(p[:, :m]), n=m, **ck)
        else:
            ck = dict(okw)
            if W.points == "param01":
                ck.update(pts_offsets=np.arange(m + 1, dtype=np.int64), points=np.ascontiguousarray(p[:, :m].T))
            elif W.points:
                ck["points"] = W.points
            ref = O.integrate2_batch(ALGO[W.algo], W.functor, W.yrange, W.a, W.b, fparams=np.ascontiguousarray(p[:, :m]),
                                     yparams=W.yparams, n=m, **ck)
        cpu_s = time.perf_counter() - t0
        allow = np.abs(W.tol.to_abs(np.abs(ref["estimate"])))
        both_nan = np.isnan(g.estimate[:m]) & np.isnan(ref["estimate"])
        diff = np.where(both_nan, 0.0, np.abs(g.estimate[:m] - ref["estimate"]))
        okr = ref["status"] == 0
        fl = eng.functor(W.dim, W.functor)[3] + 8.0 + (6.0 if not (np.isfinite(W.a) and np.isfinite(W.b)) else 0.0)
        row = dict(workload=name, n=n, ms=ms, integrals_per_s=n / (ms * 1e-3), converged_frac=float((g.status == 0).mean()),
                   mean_nevals=float(g.nevals.mean()), status_hist=np.bincount(g.status, minlength=6).tolist(),
                   tflops=float(g.nevals.astype(np.float64).sum()) * fl / (ms * 1e-3) / 1e12,
                   fp64_frac=float(g.nevals.astype(np.float64).sum()) * fl / (ms * 1e-3) / peak,
                   cpu_sample=m, cpu_integrals_per_s=m / cpu_s, cpu_threads=O.num_threads(),
                   status_mismatch=int((g.status[:m] != ref["status"]).sum()),
                   nevals_mismatch=int((g.nevals[:m] != ref["nevals"]).sum()),
                   out_of_tol=int((diff[okr] > allow[okr] + 1e-300).sum()),
                   bit_equal=int(((g.estimate[:m] == ref["estimate"]) | both_nan).sum()))
        rows.append(row)
        print(json.dumps(row), flush=True)
    print("\n| workload | n | GPU ms | integrals/s | mean nevals | converged | FP64 frac | CPU integrals/s (threads) | "
          "status / nevals mismatch | out of tol | bit-equal |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['workload']} | {r['n']} | {r['ms']:.3f} | {r['integrals_per_s']:.3g} | {r['mean_nevals']:.1f} | "
              f"{r['converged_frac']:.5f} | {r['fp64_frac']:.3f} | {r['cpu_integrals_per_s']:.3g} ({r['cpu_threads']}) | "
              f"{r['status_mismatch']} / {r['nevals_mismatch']} of {r['cpu_sample']} | {r['out_of_tol']} | {r['bit_equal']} |")
    print(f"\nFP64 peak measured in this run: {peak / 1e12:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
