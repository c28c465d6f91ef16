"""Probe: can this box's GPUs take the engine's output stores directly over NVLink (symmetric memory)?

    torchrun --nproc-per-node 2 tools/symm_probe.py
Every rank integrates its shard and the kernels write the results straight into RANK 0's buffer
(peer mapping from torch.distributed._symmetric_memory); rank 0 then checks all shards.
"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]
import gkquad_b200 as gk  # noqa: E402
from gkquad_b200.workloads import WORKLOADS  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
import torch.distributed._symmetric_memory as symm  # noqa: E402

n = 1 << 20
buf = symm.empty((world, 3 * n), dtype=torch.float64, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
root = hdl.get_buffer(0, (world, 3 * n), torch.float64)  # rank 0's buffer, mapped into this process
print(f"rank {rank}: rendezvous ok, multicast {hdl.has_multicast_support}", file=sys.stderr)

W = WORKLOADS["S1"]
eng = gk.Engine(local)
p = torch.as_tensor(W.make_params(rank * n, n), device=dev)
row = root[rank]
ints = row[2 * n:].view(torch.int32)
out = (row[:n], row[n:2 * n], ints[:n], ints[n:])
plan = eng.plan_batch(W.functor, W.a, W.b, p, algo=1, tol=W.tol, n=n, out=out)
loc = eng.integrate_batch(W.functor, W.a, W.b, p, algo=1, tol=W.tol, n=n)  # local reference
torch.cuda.synchronize()
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
hdl.barrier()
dist.barrier()
t0 = time.perf_counter()
K = 32
for _ in range(K):
    plan.run()
torch.cuda.synchronize()
hdl.barrier()
dt = (time.perf_counter() - t0) / K
# every rank sends its local result to rank 0 for the check (NCCL), rank 0 compares with what landed
chk = torch.stack([loc.estimate, loc.delta])
allchk = [torch.empty_like(chk) for _ in range(world)] if rank == 0 else None
dist.gather(chk, allchk, dst=0)
if rank == 0:
    ok = all(torch.equal(buf[r, :n], allchk[r][0]) and torch.equal(buf[r, n:2 * n], allchk[r][1]) for r in range(world))
    print(f"p2p-store gather to rank 0: {'OK' if ok else 'MISMATCH'}; {dt * 1e3:.3f} ms per step, "
          f"{world * n / dt:.3e} integrals/s over {world} GPUs")
dist.destroy_process_group()
