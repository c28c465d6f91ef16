"""Prints the per-chunk timeline of the host entry (GKQ_HOST_TRACE=1) for the S1 batch."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

os.environ["GKQ_HOST_TRACE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]
import gkquad_b200 as gk  # noqa: E402
from gkquad_b200.workloads import WORKLOADS  # noqa: E402

W = WORKLOADS["S1"]
n = 1 << 20
eng = gk.Engine(0)
p = torch.as_tensor(W.make_params(0, n)).pin_memory()
out = (torch.empty(n, dtype=torch.float64).pin_memory().numpy(), torch.empty(n, dtype=torch.float64).pin_memory().numpy(),
       torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
       torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32))
for k in range(4):
    t0 = time.perf_counter()
    eng.integrate_batch(W.functor, 0.0, 1.0, p.numpy(), algo=1, tol=W.tol, n=n, out=out)
    print(f"call {k}: {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
