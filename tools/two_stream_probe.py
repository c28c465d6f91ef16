"""Throughput of the S1 step when independent batches are driven through TWO handles on TWO streams
(the single-wave kernels of one batch fill the ramps / tails of the other) vs one handle."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]
import gkquad_b200 as gk  # noqa: E402
from gkquad_b200.workloads import WORKLOADS  # noqa: E402

W = WORKLOADS["S1"]
n = 1 << 20
NS = 8
NE = 4
engs = [gk.Engine(0) for _ in range(NE)]
streams = [torch.cuda.Stream() for _ in range(NE)]
ps = [torch.as_tensor(W.make_params(k * n, n), device="cuda:0") for k in range(NS)]
outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda"),
         torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")) for _ in range(NS)]
for e in engs:
    e.set_deferred_join(True)
plans = [[e.plan_batch(W.functor, 0.0, 1.0, ps[k], algo=1, tol=W.tol, n=n, out=outs[k]) for k in range(NS)] for e in engs]


def run(K, nstreams):
    def body(K):
        for k in range(K):
            s = k % nstreams
            with torch.cuda.stream(streams[s]):
                plans[s][k % NS].run()
        for s in range(nstreams):
            with torch.cuda.stream(streams[s]):
                engs[s].join()
    body(16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(nstreams):
        streams[s].wait_event(e0)
    t0 = time.perf_counter()
    body(K)
    t1 = time.perf_counter()
    for s in range(nstreams):
        torch.cuda.current_stream().wait_stream(streams[s])
    e1.record()
    torch.cuda.synchronize()
    print(f"streams={nstreams}: GPU {e0.elapsed_time(e1) / K * 1e3:.1f} us/step, host enqueue {(t1 - t0) / K * 1e6:.1f} us/step")


for ns in (1, 2, 3, 4, 1, 2, 3, 4):
    run(64, ns)
