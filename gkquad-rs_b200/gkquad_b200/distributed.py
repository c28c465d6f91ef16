"""Multi-GPU plumbing: batches of independent integrals shard statically across the ranks of one
node (one process per GPU) and only the results are exchanged.  There is no data-path collective
during the computation.

The product path is inside the library: `init_comm` joins every rank's Engine to an NCCL
communicator through the C ABI (gkq_comm_get_unique_id / gkq_comm_init; torch.distributed only
carries the 128-byte id) and `Engine.gather` = gkq_gather moves 20 bytes per integral over NVLink
and un-permutes them on the device.  `gather_results` below is the same exchange written with
torch.distributed collectives (24 bytes per integral); with CPU tensors it runs over gloo and is how
the host-side shard / un-permute logic is tested without GPUs (tests/test_distributed_cpu.py)."""
from __future__ import annotations

import numpy as np

from .workloads import block_cyclic_shard

BLOCK = 4096


def shard_indices(n_total: int, rank: int, world: int, block: int = BLOCK) -> np.ndarray:
    """Global indices owned by `rank`: block-cyclic, i -> rank (i / block) mod world."""
    return block_cyclic_shard(n_total, rank, world, block)


def shard_sizes(n_total: int, world: int, block: int = BLOCK):
    return [int(len(block_cyclic_shard(n_total, r, world, block))) for r in range(world)]


def pack_results(est, dlt, nev, st, n_pad: int):
    """Pack one rank's SoA results into a single float64 buffer of 3*n_pad elements
    (all ranks use the same n_pad = largest shard so that all_gather_into_tensor applies)."""
    import torch
    n = int(est.shape[0])
    buf = torch.zeros(3 * n_pad, dtype=torch.float64, device=est.device)
    buf[:n] = est
    buf[n_pad:n_pad + n] = dlt
    ints = buf[2 * n_pad:].view(torch.int32)
    ints[:n] = nev.view(torch.int32) if nev.dtype != torch.int32 else nev
    ints[n_pad:n_pad + n] = st.view(torch.int32) if st.dtype != torch.int32 else st
    return buf


def gather_results(est, dlt, nev, st, n_total: int, group=None, block: int = BLOCK):
    """All-gather the per-rank results and return (estimate, delta, nevals, status) in GLOBAL
    integral order on every rank (torch tensors on the input device)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world, block)
    n_pad = max(sizes)
    packed = pack_results(est, dlt, nev, st, n_pad)
    gathered = torch.empty(world * 3 * n_pad, dtype=torch.float64, device=est.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    return unpack_global(gathered, n_total, world, n_pad, block)


def unpack_global(gathered, n_total: int, world: int, n_pad: int, block: int = BLOCK):
    """Un-permute the rank-major gather buffer back to global integral order."""
    import torch
    dev = gathered.device
    est = torch.empty(n_total, dtype=torch.float64, device=dev)
    dlt = torch.empty(n_total, dtype=torch.float64, device=dev)
    nev = torch.empty(n_total, dtype=torch.int32, device=dev)
    st = torch.empty(n_total, dtype=torch.int32, device=dev)
    for r in range(world):
        idx = torch.as_tensor(shard_indices(n_total, r, world, block), device=dev)
        m = int(idx.shape[0])
        base = gathered[r * 3 * n_pad:(r + 1) * 3 * n_pad]
        est[idx] = base[:m]
        dlt[idx] = base[n_pad:n_pad + m]
        ints = base[2 * n_pad:].view(torch.int32)
        nev[idx] = ints[:m]
        st[idx] = ints[n_pad:n_pad + m]
    return est, dlt, nev, st


def init_comm(engine, group=None):
    """Join `engine` (this rank's handle) to the library's own NCCL communicator: rank 0 creates the
    id, torch.distributed broadcasts it, every rank calls gkq_comm_init (collective)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        engine.comm_init(1, 0, None)
        return
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(world, rank, box[0])
