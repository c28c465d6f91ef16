"""BatchPipeline: throughput mode for a stream of independent batches.

One batch of the QAGS path is a short chain of kernels (17-point pass over everything, 25-point pass
and first bisection step over the compacted survivors, one persistent launch for the stragglers).
The survivor kernels are single-wave launches: their ramp, their tail and the SMs that drew one
block fewer leave a fifth of the machine idle while they run.  Batches are independent, so the
cure is to keep two of them in flight: `depth` handles (a handle owns its scratch, worklists and
straggler records, and is not shared between streams -- include/gkquad_b200.h, "Threading") on
`depth` CUDA streams, batch k on lane k % depth, every lane in deferred-join mode.  The kernels of
one batch then fill the holes of the other (measured on B200, workload S1, 1 M integrals per batch:
142 -> 123 us per batch at depth 2, 118 us at depth 3; tools/bench_depth_sweep.sh, tools/two_stream_probe.py).

This is host-side orchestration only (the rayon-style "several Integrators at work" of the
reference's users); every lane runs exactly the calls a single Engine would.
"""
from __future__ import annotations

from .engine import Engine


class BatchPipeline:
    def __init__(self, device: int = 0, depth: int = 2, express: bool = False):
        """`express`: one more lane (index `express_lane`) whose stream and handle-owned streams have
        CUDA's high priority (gkq_create_prio): a batch run there leaves the GPU first although the
        other lanes are busy -- for the batch whose results have to travel (a gather to another GPU)
        while the rest of the pipeline still computes."""
        import torch
        assert depth >= 1
        self.device = int(device)
        self._depth = depth
        self.engines = [Engine(device) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(depth)]
        self.express_lane = None
        if express:
            self.express_lane = depth
            self.engines.append(Engine(device, stream_priority=-1))
            self.streams.append(torch.cuda.Stream(device=device, priority=-1))
        self._raw = [s.cuda_stream for s in self.streams]  # cudaStream_t handles for the C ABI
        for e in self.engines:
            e.set_deferred_join(True)

    @property
    def depth(self):
        """Number of regular lanes (the express lane is not counted)."""
        return self._depth

    def plan_batch(self, lane: int, *args, **kw):
        """Marshal a batch call on lane `lane` (see Engine.plan_batch); run it with run()."""
        plan = self.engines[lane].plan_batch(*args, **kw)
        return plan

    def run(self, plan, lane: int):
        """Enqueue a plan made by plan_batch(lane, ...) on its lane's stream.  Its outputs are complete
        after join() (or once the lane's stream and the lane's Engine.join() are)."""
        plan.run(self._raw[lane])

    def fork(self):
        """Order every lane behind the work already queued on the current stream."""
        import torch
        ev = torch.cuda.Event()
        ev.record()
        for s in self.streams:
            s.wait_event(ev)

    def join(self):
        """Finish the deferred remainders of every lane and order the current stream behind all lanes."""
        import torch
        for e, raw in zip(self.engines, self._raw):
            e._check(e.L.gkq_join(e.h, raw))
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def kernel_launches(self):
        return sum(e.stats()["kernel_launches"] for e in self.engines)

    def reset_stats(self):
        for e in self.engines:
            e.reset_stats()

    def close(self):
        import torch
        torch.cuda.synchronize(self.device)
        for e in self.engines:
            e.set_deferred_join(False)
            e.close()
