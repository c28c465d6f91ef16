"""Time of one gkq_gather (the library's NCCL path) per mode, under torchrun: 1M integrals per rank.
    python -m torch.distributed.run --nproc-per-node N tools/gather_probe.py"""
import os, sys
from pathlib import Path
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gkquad-rs_b200")]
import gkquad_b200 as gk
from gkquad_b200 import distributed as D
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))
eng = gk.Engine(lr)
D.init_comm(eng)
n = 1 << 20
res = (torch.rand(n, dtype=torch.float64, device=dev), torch.rand(n, dtype=torch.float64, device=dev),
       torch.randint(0, 2000, (n,), dtype=torch.int32, device=dev), torch.randint(0, 6, (n,), dtype=torch.int32, device=dev))
total = n * world
out = tuple(torch.empty(total, dtype=t.dtype, device=dev) for t in res)
for root in (0, -1):
    for _ in range(5):
        eng.gather(res, total, root=root, out=out)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.gather(res, total, root=root, out=out)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world} root {root}: {t.item() * 1e3:.1f} us per gather of {n} integrals per rank "
              f"({20 * n * (world - 1) / (t.item() * 1e-3) / 1e9:.0f} GB/s into a receiver)", flush=True)
dist.destroy_process_group()
