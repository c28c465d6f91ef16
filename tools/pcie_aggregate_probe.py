"""Aggregate host-link ceiling of the box: N ranks (one per GPU) copy pinned host buffers of the
e2e step's size concurrently -- H2D alone, D2H alone, both directions at once -- between barriers.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29577 tools/pcie_aggregate_probe.py [--pin-cores]

Prints per-rank and summed GB/s and the integrals/s the sum allows at the host entry's 36 B per
integral (16 B of parameters in, 20 B of packed results out).  `--pin-cores` binds each rank to its
own slice of the host cores first (the box reports ONE NUMA node, so there is no node to pick).
"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if "--pin-cores" in sys.argv:
    nc = os.cpu_count() or 1
    per = max(1, nc // world)
    os.sched_setaffinity(0, set(range(local * per, min(nc, (local + 1) * per))))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")

N = 1 << 20
H2D_B, D2H_B = 16 * N, 20 * N
NBUF = 4
h_in = [torch.empty(H2D_B, dtype=torch.uint8).pin_memory() for _ in range(NBUF)]
h_out = [torch.empty(D2H_B, dtype=torch.uint8).pin_memory() for _ in range(NBUF)]
d_in = [torch.empty(H2D_B, dtype=torch.uint8, device="cuda") for _ in range(NBUF)]
d_out = [torch.empty(D2H_B, dtype=torch.uint8, device="cuda") for _ in range(NBUF)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
REPS = 60


def run(mode):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(REPS):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s_in):
                d_in[r % NBUF].copy_(h_in[r % NBUF], non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s_out):
                h_out[r % NBUF].copy_(d_out[r % NBUF], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    nb = REPS * ((H2D_B if mode in ("h2d", "both") else 0) + (D2H_B if mode in ("d2h", "both") else 0))
    return nb / dt / 1e9


for mode in ("h2d", "d2h", "both"):
    run(mode)  # warm
    g = run(mode)
    t = torch.tensor([g], dtype=torch.float64)
    allg = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    if world > 1:
        dist.all_gather(allg, t)
    else:
        allg = [t]
    if rank == 0:
        v = [float(x.item()) for x in allg]
        line = f"N={world} {mode:4s}: per rank " + " ".join(f"{x:5.1f}" for x in v) + f"  sum {sum(v):6.1f} GB/s"
        if mode == "both":
            line += f"  => host-entry ceiling {sum(v) * 1e9 / 36.0:.3g} integrals/s at 36 B per integral"
        print(line, flush=True)
if "--sizes" in sys.argv and rank == 0 and world == 1:
    # link efficiency against copy size: both directions at once, copies of `mb` MB back to back on one
    # stream per direction (the host entry moves a chunk's arrays as copies of 0.5 - 8 MB)
    for mb in (0.5, 1, 2, 4, 8, 16):
        nb = int(mb * (1 << 20))
        per = max(1, (16 << 20) // nb)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            t0 = time.perf_counter()
            for r in range(20):
                for c in range(per):
                    with torch.cuda.stream(s_in):
                        d_in[r % NBUF][c * nb:(c + 1) * nb].copy_(h_in[r % NBUF][c * nb:(c + 1) * nb], non_blocking=True)
                    with torch.cuda.stream(s_out):
                        h_out[r % NBUF][c * nb:(c + 1) * nb].copy_(d_out[r % NBUF][c * nb:(c + 1) * nb], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = max(best, 2 * 20 * per * nb / dt / 1e9)
        print(f"both directions, copies of {mb:4} MB: {best:5.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
