"""World-size-2 gloo test of the N>1 path's host logic: block-cyclic sharding + the single
result gather + un-permute must reassemble exactly what one process computes.  The per-rank
"compute" here is the CPU oracle (test infrastructure); on GPUs the engine takes its place and
the gather runs over NCCL (bench.py --gpus N)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    for p in (ROOT, ROOT / "gkquad-rs_b200"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from gkquad_b200 import distributed as D
    from gkquad_b200.workloads import WORKLOADS
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W = WORKLOADS["S2"]
    idx = D.shard_indices(n_total, rank, world)
    p_all = W.make_params(0, n_total)
    p = np.ascontiguousarray(p_all[:, idx])
    r = O.integrate_batch(O.ALGO_QAGS, W.functor, W.a, W.b, params=p, tol=("rel", 1e-10), n=len(idx), nthreads=1)
    est, dlt, nev, st = D.gather_results(torch.as_tensor(r["estimate"]), torch.as_tensor(r["delta"]),
                                         torch.as_tensor(r["nevals"].astype(np.int32)),
                                         torch.as_tensor(r["status"].astype(np.int32)), n_total)
    np.savez(Path(out_dir) / f"rank{rank}.npz", est=est.numpy(), dlt=dlt.numpy(), nev=nev.numpy(), st=st.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_shard_and_gather_world2(tmp_path):
    import torch.multiprocessing as mp
    from gkquad_b200.workloads import WORKLOADS
    from oracle import oracle as O
    n_total = 3 * 4096 + 777  # ragged: ranks own 2 blocks / 1 block + tail
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    W = WORKLOADS["S2"]
    ref = O.integrate_batch(O.ALGO_QAGS, W.functor, W.a, W.b, params=W.make_params(0, n_total),
                            tol=("rel", 1e-10), n=n_total, nthreads=2)
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["est"], ref["estimate"]) and np.array_equal(z["dlt"], ref["delta"])
        assert np.array_equal(z["nev"].astype(np.uint32), ref["nevals"])
        assert np.array_equal(z["st"].astype(np.uint32), ref["status"])
