"""PCIe copy time versus size (pinned host memory), single copies and groups of 4 on one stream."""
import torch

torch.cuda.init()
s = torch.cuda.Stream()
for direction in ("d2h", "h2d"):
    for kb in (64, 256, 512, 1024, 2048, 4096, 8192):
        nb = kb * 1024
        h = torch.empty(nb, dtype=torch.uint8).pin_memory()
        d = torch.empty(nb, dtype=torch.uint8, device="cuda")
        hs = [torch.empty(nb, dtype=torch.uint8).pin_memory() for _ in range(4)]
        ds = [torch.empty(nb, dtype=torch.uint8, device="cuda") for _ in range(4)]
        res = []
        for group in (1, 4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for rep in range(8):
                with torch.cuda.stream(s):
                    e0.record()
                    for g in range(group):
                        if direction == "d2h":
                            hs[g].copy_(ds[g], non_blocking=True)
                        else:
                            ds[g].copy_(hs[g], non_blocking=True)
                    e1.record()
                s.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res.append(best * 1e3)
        print(f"{direction} {kb:5d} KB: 1 copy {res[0]:7.1f} us ({nb / res[0] / 1e3:5.1f} GB/s)   4 copies {res[1]:7.1f} us ({4 * nb / res[1] / 1e3:5.1f} GB/s)")
