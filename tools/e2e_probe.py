"""Breaks the end-to-end (host buffers) step into its parts on the GPU box: raw PCIe copies,
the device-resident compute, and the host entry at several pipeline chunk sizes."""
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import gkquad_b200 as gk
    from gkquad_b200.workloads import WORKLOADS
    W = WORKLOADS["S1"]
    n = 1 << 20
    eng = gk.Engine(0)
    p = torch.as_tensor(W.make_params(0, n)).pin_memory()
    out = (torch.empty(n, dtype=torch.float64).pin_memory().numpy(), torch.empty(n, dtype=torch.float64).pin_memory().numpy(),
           torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
           torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32))
    for _ in range(3):
        eng.integrate_batch(W.functor, 0.0, 1.0, p.numpy(), algo=1, tol=W.tol, n=n, out=out)
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        eng.integrate_batch(W.functor, 0.0, 1.0, p.numpy(), algo=1, tol=W.tol, n=n, out=out)
    dt = (time.perf_counter() - t0) / K
    print(f"chunk {os.environ.get('GKQ_HOST_CHUNK', 'default'):>8}: host entry {dt * 1e3:.3f} ms/step -> {n / dt:.3e} integrals/s")
    sys.exit(0)

n = 1 << 20
h_in = torch.empty(2 * n, dtype=torch.float64).pin_memory()
d_in = torch.empty(2 * n, dtype=torch.float64, device="cuda")
d_out = torch.empty(3 * n, dtype=torch.float64, device="cuda")
h_out = torch.empty(3 * n, dtype=torch.float64).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = timeit(lambda: d_in.copy_(h_in, non_blocking=True))
print(f"H2D 16.8 MB pinned: {t * 1e3:.3f} ms  ({h_in.nbytes / t / 1e9:.1f} GB/s)")
t = timeit(lambda: h_out.copy_(d_out, non_blocking=True))
print(f"D2H 25.2 MB pinned: {t * 1e3:.3f} ms  ({h_out.nbytes / t / 1e9:.1f} GB/s)")


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


t = timeit(both)
print(f"H2D + D2H concurrently: {t * 1e3:.3f} ms")
for ch in ("default", "32768", "65536", "131072", "262144", "524288", "1048576"):
    env = dict(os.environ)
    if ch != "default":
        env["GKQ_HOST_CHUNK"] = ch
    subprocess.run([sys.executable, __file__, "child"], env=env, check=True)
