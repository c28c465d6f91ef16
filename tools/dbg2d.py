"""Debug aid: the nested 2-D mappings (GKQ_NESTED_MODE 1 / 2 / 3) against the oracle on mixed-budget batches."""
import os, sys, time
import numpy as np
sys.path[:0] = ["/root/repo", "/root/repo/gkquad-rs_b200", "/root/repo/tests"]
import gkquad_b200 as gk
from gkquad_b200 import Tolerance
from oracle import oracle as O
eng = gk.Engine(0)
import torch
for n, choices, poolr in [(3000, [100000], None), (3000, [300, 100000], None), (3000, [2000, 100000], None), (3000, [20000, 100000], None),
                          (3000, [60000, 100000], None), (3000, [300, 2000, 20000, 60000, 100000], "1"), (3000, [300, 2000, 20000, 60000, 100000], "6"),
                          (96, [100000], None), (512, [2000, 100000], None)]:
    rng = np.random.default_rng(3)
    p = np.stack([-0.5 + rng.random(n), -0.5 + rng.random(n)])
    off = np.arange(n + 1, dtype=np.int64)
    budgets = rng.choice(choices, n).astype(np.uint64)
    ref = O.integrate2_batch(2, "d1_p", "const", -1.0, 1.0, fparams=p, yparams=[-1.0, 1.0], points=np.ascontiguousarray(p.T), pts_offsets=off,
                             tol=("rel", 1e-6), max_evals=100000, n=n, max_evals_each=budgets)
    line = f"n {n} budgets {choices} pool_r {poolr}:"
    for mode in ("2", "3"):
        os.environ["GKQ_NESTED_MODE"] = mode
        if poolr:
            os.environ["GKQ_POOL_R"] = poolr
        else:
            os.environ.pop("GKQ_POOL_R", None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = eng.integrate2_batch("d1_p", "const", -1.0, 1.0, fparams=p, yparams=[-1.0, 1.0], algo=2, tol=Tolerance.Relative(1e-6),
                                 max_evals=100000, pts_offsets=off, pts_xy=np.ascontiguousarray(p.T), max_evals_each=budgets, n=n)
        dt = time.perf_counter() - t0
        bad = np.nonzero(~((g.estimate == ref["estimate"]) | (np.isnan(g.estimate) & np.isnan(ref["estimate"]))) | (g.nevals != ref["nevals"]) | (g.status != ref["status"]))[0]
        line += f"  mode {mode}: {len(bad)} mismatches {np.unique(budgets[bad]).tolist()} ({dt * 1e3:.1f} ms host)"
    print(line, flush=True)
