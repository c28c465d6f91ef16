#!/usr/bin/env python
"""bench.py -- converged integrals/s (QAGS, rel_tol 1e-10) of the batched quadrature hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one pass of the hot path over one batch of synthetic parameterised integrands.
Headline (top-level keys) = BASELINE.json configs[1], workload S1: 1M smooth integrands
exp(-p0 x) cos(p1 x) on [0,1] per GPU, QAGS, Tolerance::Relative(1e-10), max_evals 2000.  Weak
scaling: every rank owns a block-cyclic shard of N x batch integrals; results stay sharded, the job
ends with one gather to rank 0 through the library's own NCCL path (gkq_gather_range, chunked behind
the kernels of the last batch).

The `configs` block of the same JSON line carries every BASELINE.json configuration at full size
(cfg-1 `simple` CPU row; cfg-3 P0 + P1; cfg-4 I0 + I1; cfg-5 D0 + D1 at 100k, block-cyclic sharded
across the ranks and gathered every step), each with value, ms_per_step, roofline, e2e, cpu_baseline
and parity_vs_cpu computed on the outputs the TIMED region wrote.

Prints ONE JSON line on rank 0 (DESIGN.md section 7 describes every field).
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "gkquad-rs_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

METRIC = "converged integrals/sec (QAGS, rel_tol 1e-10)"
UNIT = "integrals/s"
ALGO = {"AUTO": 0, "QAGS": 1, "QAGP": 2}
SHARD_BLOCK = 4096
# (name in `configs`, workload, integrals: per GPU for 1-D / in total for 2-D)
SECONDARY = [("cfg3_P0", "P0", 1 << 20), ("cfg3_P1", "P1", 1 << 20), ("cfg4_I0", "I0", 1 << 20),
             ("cfg4_I1", "I1", 1 << 20), ("cfg5_D0", "D0", 100_000), ("cfg5_D1", "D1", 100_000)]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="S1")
    ap.add_argument("--batch", type=int, default=1 << 20, help="integrals per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="all", help="'all', 'none' or a comma list of names of the `configs` block")
    ap.add_argument("--e2e-threads", type=int, default=4,
                    help="host threads (one handle each) that make the blocking host-entry calls of the e2e leg")
    ap.add_argument("--pipeline-depth", type=int, default=3,
                    help="handles / streams the independent batches alternate between (1 = one handle, one stream)")
    ap.add_argument("--final-lane", default="express", choices=["express", "regular"],
                    help="N > 1: the job's last batch (the one that is gathered) runs on a high-priority lane of the "
                         "pipeline, so its transfer overlaps what the other lanes still compute; 'regular' = a normal lane")
    ap.add_argument("--start-barrier", default="device", choices=["device", "host"],
                    help="N > 1: after barrier + synchronize the ranks also meet in a one-element all-reduce on the "
                         "stream right before the start event (their clocks start together); 'host' = host barrier only")
    ap.add_argument("--force-express", action="store_true",
                    help="diagnostics at N = 1: run the last batch on the express lane as the N > 1 path does (no gather)")
    ap.add_argument("--no-pin", action="store_true", help="N > 1: do not bind each rank to its own slice of the host cores")
    ap.add_argument("--gather-chunks", type=int, default=1,
                    help="N > 1: the last batch is computed and gathered in this many chunks (only the last chunk's transfer is exposed)")
    return ap.parse_args()


def workload_config(W, n, world):
    """The `config` object: identical in both arms (ours / --impl reference) for the same N."""
    return {"workload": f"{W.name}: {W.description}; {W.algo}, {W.tol}, max_evals {W.max_evals}",
            "batch_per_gpu": n, "global_batch": n * world, "parallelism": f"shard{world}",
            "l2": "8 independent batches cycled (320 MB of inputs+outputs > 126 MB L2); no flush needed"}


def shard_params(W, batch, rank, world, offset=0, total=None):
    """Block-cyclic shard of the global batch; every rank regenerates its own
    parameters from the counter-based generator (`offset` shifts the counters: another batch of the
    same family).  total = global batch size (default world * batch)."""
    from gkquad_b200.workloads import block_cyclic_shard
    idx = block_cyclic_shard(batch * world if total is None else total, rank, world, SHARD_BLOCK)
    blocks = []
    B = SHARD_BLOCK
    for s in range(0, len(idx), B):  # the generator is indexed by the global integral number
        blocks.append(W.make_params(int(idx[s]) + offset, min(B, len(idx) - s)))
    if not blocks:
        return W.make_params(0, 0), idx
    return np.ascontiguousarray(np.concatenate(blocks, axis=1)), idx


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        """Block until nvidia-smi has attached to the driver and delivered its first row."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        """Rows that arrived between t_begin and t_end + one sampling period (perf_counter values of
        the timed region); all rows when the region was too short to catch one."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t_begin is None or t_begin <= t <= t_end + 0.12]
        if not rows:
            rows = [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def tol_to_oracle(t):
    return {0: ("abs", t.abs_tol), 1: ("rel", t.rel_tol), 2: ("abs_or_rel", t.abs_tol, t.rel_tol),
            3: ("abs_and_rel", t.abs_tol, t.rel_tol)}[t.kind]


# ---- the reference's CPU implementation of the path = the golden-vector-pinned oracle port ---------
def oracle_call(W, p, n, nthreads=0):
    """One pass of the oracle over a batch of workload W (1-D or 2-D); returns the result dict."""
    from oracle import oracle as O
    kw = dict(tol=tol_to_oracle(W.tol), max_evals=W.max_evals, n=n, nthreads=nthreads)
    if W.dim == 1:
        if W.points == "param0":
            kw.update(pts_offsets=np.arange(n + 1, dtype=np.int64), points=np.ascontiguousarray(p[0]))
        elif W.points:
            kw["points"] = W.points
        return O.integrate_batch(ALGO[W.algo], W.functor, W.a, W.b, params=p, **kw)
    if W.points == "param01":
        kw.update(pts_offsets=np.arange(n + 1, dtype=np.int64), points=np.ascontiguousarray(p.T))
    elif W.points:
        kw["points"] = W.points
    return O.integrate2_batch(ALGO[W.algo], W.functor, W.yrange, W.a, W.b, fparams=p, yparams=W.yparams, **kw)


def cpu_leg(W, p, n, seconds_target=6.0, min_reps=3, max_reps=50):
    """The oracle on all host threads over the same batch: (converged integrals/s of the best pass,
    threads, passes, last result)."""
    from oracle import oracle as O
    m = min(n, 65536)
    oracle_call(W, np.ascontiguousarray(p[:, :m]), m)  # warm-up (page faults, thread pool)
    reps, t0, best = 0, time.perf_counter(), 1e30
    while True:
        t1 = time.perf_counter()
        r = oracle_call(W, p, n)
        best = min(best, time.perf_counter() - t1)
        reps += 1
        if (reps >= min_reps and time.perf_counter() - t0 > seconds_target) or reps >= max_reps:
            break
    return int((r["status"] == 0).sum()) / best, O.num_threads(), reps, r


def parity(W, g_est, g_nev, g_st, ref):
    """Mismatch counts of engine outputs against the oracle on the same integrals."""
    est = ref["estimate"]
    tol_abs = np.abs(W.tol.to_abs(np.abs(est)))
    ok = ref["status"] == 0
    both_nan = np.isnan(g_est) & np.isnan(est)
    diff = np.where(both_nan, 0.0, np.abs(g_est - est))
    return {"n": int(len(est)), "status_mismatch": int((g_st != ref["status"]).sum()),
            "nevals_mismatch": int((g_nev != ref["nevals"]).sum()),
            "out_of_tolerance": int((diff[ok] > tol_abs[ok] + 1e-300).sum()),
            "bit_equal_estimates": int(((g_est == est) | both_nan).sum())}


def cfg1_simple():
    """BASELINE configs[0]: the reference's own CPU-runnable case -- `simple`: x^2 on [0,1] with the
    Integrator defaults (benches/single.rs:10-17; README.md:45 quotes 88.351 ns per call on an
    i5-8265U).  The oracle's ns/call on ONE thread of this host, in a compiled loop."""
    from oracle import oracle as O
    n = 1 << 20
    t_best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        r = O.integrate_batch(0, "square", 0.0, 1.0, tol=("abs_or_rel", 1.49e-8, 1.49e-8), max_evals=2000, n=n, nthreads=1)
        t_best = min(t_best, time.perf_counter() - t0)
    one = O.integrate_one(0, "square", 0.0, 1.0)
    return {"workload": "simple: x^2 on [0,1], Integrator defaults (AUTO -> QAGS, AbsOrRel(1.49e-8, 1.49e-8), max_evals 2000)",
            "impl": "CPU oracle port, 1 thread, compiled loop over 2^20 calls",
            "ns_per_call": t_best / n * 1e9, "reference_readme_ns_per_call": 88.351,
            "reference_readme_cpu": "i5-8265U (README.md:45)",
            "result": [one["estimate"], one["delta"], int(one["nevals"]), int(one["status"])],
            "all_calls_identical": bool((r["nevals"] == 17).all() and (r["status"] == 0).all()
                                        and (r["estimate"] == one["estimate"]).all())}


def run_reference(args, rank, world, real_stdout):
    """--impl reference: the reference's CPU path on the host cores (rank 0 only), on the global batch
    of our arm's config."""
    if rank != 0:
        return
    from gkquad_b200.workloads import WORKLOADS
    from oracle import oracle as O
    W = WORKLOADS[args.workload]
    n = args.batch * world
    p = W.make_params(0, n)
    for _ in range(max(1, min(args.warmup, 3))):
        r = oracle_call(W, p, n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = oracle_call(W, p, n)
    dt = (time.perf_counter() - t0) / args.steps
    val = int((r["status"] == 0).sum()) / dt
    cores = O.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(W, args.batch, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"the full {n}-integral global batch per step, {cores} host threads (golden-vector-pinned "
                                   "C++ port of gkquad's CPU path; gkquad itself needs a Rust toolchain, absent here)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(real_stdout, line)


def main():
    args = parse()
    # libraries (NCCL banner, ...) must not write to stdout: the driver reads ONE JSON line from it
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main(args, real_stdout)
    finally:
        os.dup2(real_stdout, 1)


def emit(real_stdout, line):
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def ncu_traffic(kernel_regex):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel whose ncu name matches
    `kernel_regex`, read from the newest committed `ncu --set full` raw CSV under profiles/ that has
    it (bytes; None when no capture of that kernel is committed)."""
    import re
    best = None
    for f in sorted((ROOT / "profiles").glob("ncu_*_raw.csv"), key=lambda q: q.name, reverse=True):
        try:
            rows = list(csv.reader(open(f)))
            h = rows[0]
            ik, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            units = rows[1]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for row in rows[2:]:
                if re.search(kernel_regex, row[ik]):
                    best = {"bytes": float(row[ir]) * scale.get(units[ir], 1.0) + float(row[iw]) * scale.get(units[iw], 1.0),
                            "read_bytes": float(row[ir]) * scale.get(units[ir], 1.0), "source": f"profiles/{f.name}"}
                    for key, col in (("registers_per_thread", "launch__registers_per_thread"),
                                     ("warps_active_pct_of_peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                                     ("fp64_pipe_busy_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                                     ("inst_per_cycle_per_sm", "sm__inst_executed.avg.per_cycle_elapsed")):
                        if col in h:
                            try:
                                best[key] = float(row[h.index(col)])
                            except ValueError:
                                pass
                    break
        except Exception:
            continue
        if best:
            return best
    return None


def _main(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        # (torchrun pins OMP_NUM_THREADS = 1; the oracle sizes its own std::thread pool from the hardware)
        run_reference(args, rank, max(world, args.gpus), real_stdout)
        return

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device: there is no CPU fallback"
    pinned_cores = None
    if world > 1 and not args.no_pin:
        # One slice of the host cores per rank (the boxes of this pool report ONE NUMA node, so there is no
        # node to choose): a rank whose enqueueing thread is descheduled for a millisecond starves its GPU
        # inside a 2.5 ms region and rank 0 then waits for its results (details.per_rank shows it).
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local_rank * per:(local_rank + 1) * per] or cores
            os.sched_setaffinity(0, set(mine))
            pinned_cores = [mine[0], mine[-1]]
        except Exception:
            pinned_cores = None
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    import gkquad_b200 as gk
    from gkquad_b200 import distributed as D
    from gkquad_b200.workloads import WORKLOADS

    W = WORKLOADS[args.workload]
    assert W.dim == 1, "the headline is a 1-D workload; the 2-D configurations run in the `configs` block"
    n = args.batch
    algo = ALGO[W.algo]
    # throughput mode: `depth` handles on `depth` streams, batch k on lane k % depth (BatchPipeline);
    # lane 0's handle also serves the per-kernel profiling pass, the host entry (e2e) and the gathers
    EXPRESS = (world > 1 or args.force_express) and args.final_lane == "express"
    pipe = gk.BatchPipeline(local_rank, depth=args.pipeline_depth, express=EXPRESS)
    DEPTH = pipe.depth
    eng = pipe.engines[0]
    D.init_comm(eng)  # gkq_comm_init: the library's own NCCL communicator (torch only carries the id)

    # 2 groups x G independent batches are cycled through, so that a step never finds its inputs or
    # outputs in the 126 MB L2 (8 x (16 MB in + 24 MB out) = 320 MB): "inputs larger than L2".
    G, NGROUPS = 4, 2
    NSETS = G * NGROUPS
    sets = []

    def plan_kw(p_host):
        kw = dict(algo=algo, tol=W.tol, max_evals=W.max_evals)
        if W.points and W.points != "param0":
            kw["points"] = W.points
        elif W.points == "param0":
            from gkquad_b200.workloads import csr_points_from_param0
            off, pts = csr_points_from_param0(p_host)
            kw["pts_offsets"] = torch.as_tensor(off, device=dev)
            kw["pts"] = torch.as_tensor(pts, device=dev)
        return kw

    for s in range(NSETS):
        p_host, idx = shard_params(W, n, rank, world, offset=s * n * world)
        kw = dict(plan_kw(p_host), n=n)
        out = (torch.zeros(n, dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev),
               torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev))
        p_dev = torch.as_tensor(p_host, device=dev)
        # the call is marshalled once (an Integrator owns integrand + config and is run many times)
        plan = pipe.plan_batch(s % DEPTH, W.functor, W.a, W.b, p_dev, out=out, **kw)
        sets.append(dict(p_host=p_host, p_dev=p_dev, kw=kw, plan=plan, out=out, lane=s % DEPTH))
    p_host = sets[0]["p_host"]
    comm = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
    n_total = n * world
    root_out = None
    if world > 1 and rank == 0:
        root_out = (torch.empty(n_total, dtype=torch.float64, device=dev), torch.empty(n_total, dtype=torch.float64, device=dev),
                    torch.empty(n_total, dtype=torch.int32, device=dev), torch.empty(n_total, dtype=torch.int32, device=dev))
    # the last batch of the job is computed in chunks, each gathered behind the next one's kernels
    CH = max(1, args.gather_chunks) if world > 1 else 1
    chunk_plans = {}

    def chunk_plan(s, c):
        """Plan of chunk c of set s: the same arrays, positions [lo, hi) (offsets are multiples of 4)."""
        key = (s, c)
        if key not in chunk_plans:
            S = sets[s]
            step = (n // CH + 3) // 4 * 4
            lo, hi = c * step, min(n, (c + 1) * step)
            kw = dict(S["kw"], n=hi - lo)
            if "pts_offsets" in kw:
                kw["pts_offsets"] = (kw["pts_offsets"][lo:hi + 1] - kw["pts_offsets"][lo]).contiguous()
                kw["pts"] = kw["pts"][int(S["kw"]["pts_offsets"][lo]):int(S["kw"]["pts_offsets"][hi])].contiguous()
            lane = pipe.express_lane if EXPRESS else c % DEPTH
            pl = pipe.plan_batch(lane, W.functor, W.a, W.b, S["p_dev"][:, lo:hi].contiguous(),
                                 out=tuple(t[lo:hi] for t in S["out"]), **kw)
            chunk_plans[key] = (pl, lane, lo, hi)
        return chunk_plans[key]

    def gather_set(s, stream_obj, lo=None, hi=None):
        """gkq_gather / gkq_gather_range of set s's results to rank 0 on `stream_obj`."""
        kw = {} if lo is None else dict(local_offset=lo, local_count=hi - lo)
        eng.gather(sets[s]["out"], n_total, root=0, block=SHARD_BLOCK, out=root_out, stream=stream_obj.cuda_stream, **kw)

    def step(k, per_step_gather=False):
        """One pass of the hot path over one batch: the rank integrates its shard and leaves the results
        in its own HBM.  per_step_gather (N > 1): the batch is also gathered to rank 0 on the side
        stream while the next batches are computed."""
        S = sets[k % NSETS]
        pipe.run(S["plan"], S["lane"])
        if per_step_gather:
            # the batch's stragglers are finished on its OWN lane (the lanes' latency-bound remainder
            # launches overlap one another), then the side stream gathers it
            e = pipe.engines[S["lane"]]
            e._check(e.L.gkq_join(e.h, pipe._raw[S["lane"]]))
            ev = torch.cuda.Event()
            ev.record(pipe.streams[S["lane"]])
            comm.wait_event(ev)
            gather_set(k % NSETS, comm)

    def last_step_chunked(k):
        """The job's last batch: CH chunks on alternating lanes; chunk c is gathered to rank 0 on the
        side stream (gkq_gather_range) while chunk c + 1 is integrated."""
        s = k % NSETS
        for c in range(CH):
            pl, lane, lo, hi = chunk_plan(s, c)
            pipe.run(pl, lane)
            e = pipe.engines[lane]
            e._check(e.L.gkq_join(e.h, pipe._raw[lane]))  # the chunk's stragglers on its own lane ...
            ev = torch.cuda.Event()
            ev.record(pipe.streams[lane])
            comm.wait_event(ev)
            gather_set(s, comm, lo, hi)  # ... then its gather on the side stream, behind the next chunk's kernels

    def last_step_express_only(k):
        pl, lane, lo, hi = chunk_plan(k % NSETS, 0)
        pipe.run(pl, lane)
        e = pipe.engines[lane]
        e._check(e.L.gkq_join(e.h, pipe._raw[lane]))

    ev_compute = torch.cuda.Event(enable_timing=True)

    def join(mark=False):
        pipe.join()  # outstanding deferred remainders; the current stream waits for every lane
        if mark:
            ev_compute.record()  # every lane's kernels are done; what follows is the gather's exposed part
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)

    steps = args.steps
    start_token = torch.zeros(1, device=dev)
    if world > 1:
        dist.all_reduce(start_token)  # (NCCL's lazy connection set-up happens here, not in the region)
    # nvidia-smi is started BEFORE the warm-up: while it attaches to the driver, CUDA launches of this
    # process stall for milliseconds, which must not fall into the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    # No collector pause between two launches of the region: collected and switched off BEFORE the warm-up
    # (a gc.collect() between warm-up and region idles the GPU for tens of milliseconds, its clocks drop,
    # and the first timed steps run while they ramp up again: 0.119 -> 0.124 ms per step at K = 20).
    import gc
    gc.collect()
    gc.disable()
    # (at least two deferred-join cycles of 8 batches per lane: every scratch buffer has reached its
    # steady-state size before the timed region)
    for k in range(max(18 * DEPTH, 2 * NSETS, args.warmup, 3)):
        step(k)
    join()
    if world > 1 or EXPRESS:
        for s in range(NSETS):  # chunk plans are marshalled (and their scratch sized) outside the region
            for c in range(CH):
                chunk_plan(s, c)
        for _ in range(3):  # (the express handle's remainder-grid estimate needs a flush or two behind it)
            if world > 1:
                last_step_chunked((steps - 1) % NSETS)
            else:
                last_step_express_only((steps - 1) % NSETS)
            join()
    torch.cuda.synchronize()

    # ---- timed region: EXACTLY K steps between barrier + synchronize, device-timed -------------
    # The path shards with no data-path collective: each rank's results stay in its HBM; the job ends
    # with the gather of the last batch to rank 0 (inside the region).
    pipe.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if world > 1 and args.start_barrier == "device":
        # the ranks leave the host barrier tens of microseconds apart; a one-element all-reduce on the
        # stream lines their start events up on the device (rank 0's clock would otherwise also count how
        # much later the slowest rank started, because the job ends with rank 0 receiving from everybody)
        dist.all_reduce(start_token)
    e0.record()
    pipe.fork()  # the lanes start behind the start event
    t_region0 = time.perf_counter()
    for k in range(steps - 1 if (world > 1 or EXPRESS) else steps):
        step(k)
    if world > 1:
        comm.wait_stream(torch.cuda.current_stream())
        last_step_chunked(steps - 1)
    elif EXPRESS:
        last_step_express_only(steps - 1)
    t_enq = (time.perf_counter() - t_region0) / steps
    join(mark=True)  # ... and the end event behind every lane (and the side stream)
    e1.record()
    torch.cuda.synchronize()
    t_region1 = time.perf_counter()
    gc.enable()
    if world > 1:
        dist.barrier()
    launches = pipe.kernel_launches()  # our kernels only (NCCL's are not counted)
    total_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    # diagnosis of the multi-GPU tail: per rank, the whole region, the part up to the last kernel of any
    # lane, and how long the host took to enqueue the region
    per_rank = None
    if world > 1:
        mine = torch.tensor([total_ms, e0.elapsed_time(ev_compute), (t_region1 - t_region0) * 1e3, t_enq * steps * 1e3],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"region_ms": [round(float(x[0]), 4) for x in allr], "kernels_done_ms": [round(float(x[1]), 4) for x in allr],
                    "host_wall_ms": [round(float(x[2]), 4) for x in allr], "host_enqueue_ms": [round(float(x[3]), 4) for x in allr]}
    # snapshot of what the timed region wrote (the profiling pass below re-runs set 0)
    timed_out = tuple(x.clone() for x in sets[0]["out"])
    conv = sum(float((S["out"][3] == 0).sum().item()) for S in sets) / NSETS  # mean per step
    conv_local = torch.tensor([conv], dtype=torch.float64, device=dev)
    gathered_ok = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(conv_local, op=dist.ReduceOp.SUM)
        if rank == 0:  # what arrived on rank 0 is what rank 0 itself computed for its shard
            idx0 = torch.as_tensor(D.shard_indices(n_total, 0, world, SHARD_BLOCK), device=dev)
            mine = sets[(steps - 1) % NSETS]["out"]
            gathered_ok = bool(torch.equal(root_out[0][idx0], mine[0]) and torch.equal(root_out[3][idx0], mine[3]))
    total_ms = float(t.item())
    converged = float(conv_local.item())
    ms_per_step = total_ms / steps
    value = converged / (ms_per_step * 1e-3)
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None

    # ---- N > 1 only: the same K steps with EVERY batch gathered to rank 0 (gkq_gather on a side
    # stream, 20 B per integral).  Rank 0 then receives 20 B x (N - 1) x the per-GPU rate: at N = 8 that
    # is above what one GPU's NVLink ingress carries, so this figure is link-bound, not kernel-bound.
    per_step_gather = None
    if world > 1:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in range(NSETS):
            step(k, True)
        join()
        dist.barrier()
        torch.cuda.synchronize()
        g0.record()
        pipe.fork()
        comm.wait_stream(torch.cuda.current_stream())
        for k in range(steps):
            step(k, True)
        join()
        g1.record()
        torch.cuda.synchronize()
        dist.barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        ms_g = float(tg.item()) / steps
        ingress = 20.0 * n * (world - 1) / (ms_g * 1e-3) / 1e9
        per_step_gather = {"value": converged / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g,
                           "root_ingress_GBps": ingress, "bytes_per_integral": 20,
                           "link_bound_ms_per_step": 20.0 * n * (world - 1) / 770e9 * 1e3}
    torch.cuda.synchronize()
    for e in pipe.engines:
        e.set_deferred_join(False)
    p_dev, kw = sets[0]["p_dev"], sets[0]["kw"]
    est, dlt, nev, st = sets[0]["out"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # L2 flush for the profiling pass

    # ---- per-kernel profile (separate pass, events around every kernel on the launch stream) ----
    eng.set_profiling(True)
    prof_acc = {}
    PR = 5
    for _ in range(PR):
        flush.zero_()
        eng.integrate_batch(W.functor, W.a, W.b, p_dev, out=(est, dlt, nev, st), **kw)
        for name, ms in eng.profile():
            prof_acc.setdefault(name, []).append(ms)
    eng.set_profiling(False)
    eng.sync()
    stats = eng.stats()
    nev_h = timed_out[2].cpu().numpy().astype(np.int64)
    fid = eng.functor(1, W.functor)[0]
    fl_eval = eng.functor(1, W.functor)[3]
    F_t = 6.0 if not (np.isfinite(W.a) and np.isfinite(W.b)) else 0.0
    w_eval = fl_eval + F_t + 8.0  # W_i = nevals_i * (F_f + F_t + 8)
    evals_by_kernel = kernel_evals(W, n, nev_h, stats)
    kernels = []
    for name, v in prof_acc.items():
        ms = float(np.mean(v))
        fl = evals_by_kernel.get(name, 0) * w_eval
        kernels.append({"kernel": name, "ms": ms, "evals": int(evals_by_kernel.get(name, 0)),
                        "tflops": fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0})
    peak_flops, peak_ms = eng.measure_fp64_peak(5)
    dom = max(kernels, key=lambda k: k["ms"] if k["evals"] else 0.0) if kernels else None
    whole_tflops = float(nev_h.sum()) * w_eval / (ms_per_step * 1e-3) / 1e12
    roofline = None
    if dom:
        traffic = ncu_traffic(ncu_name_regex(dom["kernel"], fid))
        roofline = {
            "bound": "fp64", "kernel": dom["kernel"], "achieved": dom["tflops"], "peak": peak_flops / 1e12,
            "unit": "TFLOP/s", "frac": dom["tflops"] / (peak_flops / 1e12),
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of that kernel, read from the
            # committed ncu --set full capture (profiles/ncu_*_raw.csv) by kernel name; bytes
            "traffic": traffic["bytes"] if traffic else None,
            "traffic_source": traffic["source"] if traffic else None,
            "algorithmic_bytes": int(n * (8 * p_host.shape[0] + 24)),
            "peak_source": "measured in this run: dependent-chain DFMA micro-kernel over all SMs (MEASURED_PEAKS.json has no FP64 figure)",
            "flops_per_eval": w_eval, "kernels": kernels,
            "whole_step": {"achieved": whole_tflops, "frac": whole_tflops / (peak_flops / 1e12)},
            # SM occupancy / pipe utilisation of that kernel, from the same committed capture
            "occupancy": ({k: traffic[k] for k in ("registers_per_thread", "warps_active_pct_of_peak", "fp64_pipe_busy_pct",
                                                   "inst_per_cycle_per_sm") if k in traffic} if traffic else None),
            # north_star's worklist compaction is not a pass of its own: survivors are appended to the next
            # kernel's worklist inside the rule kernels (ballot + popc + one atomic per warp per list)
            "compaction": {
                "fused_into": "k_qags_init<qk17> / <qk25> (u32 worklists) and k_qags_first (176-byte straggler records)",
                "survivors_per_step": {"to_qk25": int(stats["last_survivors_qk25"]),
                                       "loop_entrants": int(sum(stats["last_chain_loop"])),
                                       "straggler_records": int(stats["last_survivors_step2"])},
                "bytes_per_step": int(8 * (stats["last_survivors_qk25"] + sum(stats["last_chain_loop"]))
                                      + 32 * 2 * sum(stats["last_chain_loop"]) + 176 * 2 * stats["last_survivors_step2"]),
                "GBps_at_step_rate": (8 * (stats["last_survivors_qk25"] + sum(stats["last_chain_loop"]))
                                      + 64 * sum(stats["last_chain_loop"]) + 352 * stats["last_survivors_step2"]) / (ms_per_step * 1e-3) / 1e9,
                "hbm_peak_GBps": 6563.2,
                "note": "4 B written + 4 B read per worklist entry, 32 B + 32 B per parked loop entrant, 176 B + 176 B per "
                        "straggler record; cost of the appends measured by compiling them out: 4 % of k_qags_init<qk17> "
                        "(profiles/S1_r01.md)"},
        }

    # ---- e2e: the public host entry (gkq_integrate_batch_host) with pinned host buffers ----------
    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg_1d(args, pipe, W, sets[0]["p_host"], n, world, dev, steps)

    # ---- CPU baseline beside it (rank 0, N=1 only) + parity of the TIMED outputs ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, reps, ref = cpu_leg(W, p_host, n)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the same {n}-integral batch, best of {reps} passes, all host threads (golden-vector-pinned C++ port of gkquad's CPU path)",
               "parity_vs_cpu": dict(parity(W, timed_out[0].cpu().numpy(), nev_h, timed_out[3].cpu().numpy().astype(np.uint32), ref),
                                     checked="outputs of set 0 as written by the timed region (snapshot before the profiling pass)")}

    # ---- every other BASELINE configuration ----------------------------------------------------------
    configs = {}
    want = [] if args.configs == "none" else [c for c in SECONDARY if args.configs == "all" or c[0] in args.configs.split(",")]
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.configs != "none":
        configs["cfg1_simple"] = cfg1_simple()
    configs["cfg2_S1"] = "the top-level keys of this line"
    sec_steps = max(3, min(steps, 10))
    for cname, wname, cn in want:
        try:
            configs[cname] = run_secondary(args, eng, WORKLOADS[wname], cn, rank, world, dev, sec_steps, peak_flops, flush)
        except Exception as exc:  # a failing secondary configuration must not take the headline down
            configs[cname] = {"error": repr(exc)[:300]}
    if world > 1:
        dist.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(W, n, world),
            "details": {"pipeline_depth": DEPTH, "deferred_join": True,
                        "gather": (f"final gather of the last batch to rank 0 inside the timed region: gkq_gather_range, {CH} chunks, "
                                   "20 B/integral, NCCL send/recv in the library; the last batch runs on "
                                   + ("the pipeline's high-priority lane (its transfer overlaps the other lanes' kernels)" if EXPRESS else "a regular lane")
                                   + f"; start barrier: {args.start_barrier}") if world > 1 else "none (N=1)",
                        "gathered_matches_local": gathered_ok,
                        "converged_fraction": converged / (n * world), "mean_nevals": float(nev_h.mean()),
                        "host_enqueue_ms_per_step": t_enq * 1e3, "per_rank": per_rank,
                        "host_cores_of_rank0": pinned_cores},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "configs": configs,
        }
        if per_step_gather is not None:
            line["per_step_gather"] = per_step_gather
        emit(real_stdout, line)
    if world > 1:
        dist.destroy_process_group()


def kernel_evals(W, n, nev_h, stats):
    """Integrand evaluations attributed to each kernel of one batch call (profiling names)."""
    n25 = stats["last_survivors_qk25"]
    cl, c2 = stats["last_chain_loop"], stats["last_chain_step2"]
    total = int(nev_h.sum())
    if W.algo == "QAGS" or (W.algo == "AUTO" and not W.points):
        rest = total - 17 * n - 25 * n25 - 50 * (cl[0] + cl[1])
        return {"k_qags_init<qk17>": 17 * n, "k_qags_init<qk25>": 25 * n25,
                "k_qags_first[0]": 50 * cl[0], "k_qags_first[1]": 50 * cl[1], "k_adaptive<QAGS>": rest}
    # QAGP: the flat seeding / first-step kernels count their own share when the engine reports it
    out = {}
    seeded = stats.get("last_qagp_seed_evals", 0)
    first = stats.get("last_qagp_first_evals", 0)
    if seeded:
        out["k_qagp_seed"] = seeded
    if first:
        out["k_qagp_first"] = first
    out["k_adaptive<QAGP>"] = total - seeded - first
    return out


def ncu_name_regex(display, fid):
    """ncu kernel-name pattern of a profiling display name."""
    return {"k_qags_init<qk17>": rf"k_qags_init<F1<{fid}>, 8,", "k_qags_init<qk25>": rf"k_qags_init<F1<{fid}>, 12,",
            "k_qags_first[0]": rf"k_qags_first<F1<{fid}>", "k_qags_first[1]": rf"k_qags_first<F1<{fid}>",
            "k_adaptive<QAGS>": rf"k_adaptive<F1<{fid}>, 0,", "k_adaptive<QAGP>": rf"k_adaptive<F1<{fid}>, 1,",
            "k_qagp_seed": rf"k_qagp_seed<F1<{fid}>", "k_qagp_first": rf"k_qagp_first<F1<{fid}>"}.get(display, display.split("<")[0])


def e2e_leg_1d(args, pipe, W, p_host, n, world, dev, steps):
    """The same metric through the public host entry (gkq_integrate_batch_host): pinned host buffers,
    H2D of the step's inputs and D2H of its results inside the timed call; `--e2e-threads` host threads
    with one handle each make the blocking call in turn (ctypes releases the GIL), so the H2D of one
    call overlaps the D2H of the other."""
    import torch
    import torch.distributed as dist
    algo = ALGO[W.algo]
    ph = torch.as_tensor(p_host).pin_memory()
    kwh = dict(algo=algo, tol=W.tol, max_evals=W.max_evals, n=n)
    csr = W.points == "param0"
    if W.points and not csr:
        kwh["points"] = W.points
    if csr:
        kwh["pts_offsets"] = np.arange(n + 1, dtype=np.int64)
        kwh["pts"] = ph.numpy()[0]
    nth = max(1, args.e2e_threads)
    import gkquad_b200 as gk
    engines = list(pipe.engines[:nth])
    while len(engines) < nth:  # one handle per host thread
        engines.append(gk.Engine(torch.cuda.current_device()))
    hplans, houts = [], []
    for t in range(nth):
        bufs = (torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
                torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory())
        houts.append(bufs)
        o_t = (bufs[0].numpy(), bufs[1].numpy(), bufs[2].numpy().view(np.uint32), bufs[3].numpy().view(np.uint32))
        hplans.append(engines[t].plan_batch(W.functor, W.a, W.b, ph.numpy(), **dict(kwh, out=o_t)))
    for hp_ in hplans:
        for _ in range(3):
            hp_.run()
    if world > 1:
        dist.barrier()
    e2e_steps = steps - steps % nth if steps >= nth else nth

    def host_worker(t):
        for _ in range(e2e_steps // nth):
            hplans[t].run()
    # the region (e2e_steps blocking calls, wall clock around all of them) is run three times and the
    # MEDIAN is reported: K = 20 calls last ~10 ms, and one scheduling hiccup of a host thread doubles that
    dts = []
    for rep in range(4):  # rep 0 = untimed: the handles' first CONCURRENT calls (512 k chunk schedule, pinned staging)
        if world > 1:
            dist.barrier()
        workers = [threading.Thread(target=host_worker, args=(t,)) for t in range(1, nth)]
        t0 = time.perf_counter()
        for w_ in workers:
            w_.start()
        host_worker(0)
        for w_ in workers:
            w_.join()
        if rep:
            dts.append((time.perf_counter() - t0) / e2e_steps)
    dt = sorted(dts)[1]
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    ce = torch.tensor([float((houts[0][3].numpy() == 0).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ce, op=dist.ReduceOp.SUM)
    h2d = int(p_host.nbytes) + (8 * (n + 1) + 8 * n if csr else 0)
    return {"value": float(ce.item()) / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": int(24 * n), "ms_per_step": float(tt.item()) * 1e3,
            "region_ms_per_step": [x * 1e3 for x in dts],
            "api": "gkq_integrate_batch_host, pinned host buffers" + (f", {nth} host threads with one handle each" if nth > 1 else "")
                   + "; median of 3 runs of the region"}


def run_secondary(args, eng, W, cn, rank, world, dev, steps, peak_flops, flush):
    """One BASELINE configuration of the `configs` block at full size on the already warm engine.
    1-D: cn integrals per GPU (weak); 2-D: cn integrals in total, block-cyclic sharded over the ranks
    (strong, BASELINE configs[4]) and gathered to rank 0 by gkq_gather inside every timed step."""
    import torch
    import torch.distributed as dist
    two_d = W.dim == 2
    total = cn if two_d else cn * world
    NS = 1 if two_d else 4  # 1-D: 4 independent batches cycled (> L2); 2-D: the interval slabs alone are GBs
    sets = []
    for s in range(NS):
        p_host, idx = shard_params(W, cn, rank, world, offset=s * total, total=total)
        m = int(p_host.shape[1])
        p_dev = torch.as_tensor(p_host, device=dev)
        kw = dict(algo=ALGO[W.algo], tol=W.tol, max_evals=W.max_evals, n=m)
        if not two_d:
            if W.points == "param0":
                kw["pts_offsets"] = torch.arange(m + 1, dtype=torch.int64, device=dev)
                kw["pts"] = p_dev[0].contiguous()
            elif W.points:
                kw["points"] = W.points
            out = (torch.zeros(m, dtype=torch.float64, device=dev), torch.zeros(m, dtype=torch.float64, device=dev),
                   torch.zeros(m, dtype=torch.int32, device=dev), torch.zeros(m, dtype=torch.int32, device=dev))
            plan = eng.plan_batch(W.functor, W.a, W.b, p_dev, out=out, **kw)
            sets.append(dict(p_host=p_host, run=plan.run, out=out, m=m))
        else:
            if W.points == "param01":
                kw["pts_offsets"] = torch.arange(m + 1, dtype=torch.int64, device=dev)
                kw["pts_xy"] = p_dev.t().contiguous()
            elif W.points:
                kw["points"] = W.points
            S = dict(p_host=p_host, out=None, m=m)

            def run2(S=S, p_dev=p_dev, kw=kw):
                r = eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=p_dev, yparams=W.yparams, **kw)
                S["out"] = (r.estimate, r.delta, r.nevals, r.status)
                S["keep"] = r
                return r
            S["run"] = run2
            sets.append(S)
    g_out = None
    if two_d and world > 1 and rank == 0:
        g_out = (torch.empty(total, dtype=torch.float64, device=dev), torch.empty(total, dtype=torch.float64, device=dev),
                 torch.empty(total, dtype=torch.int32, device=dev), torch.empty(total, dtype=torch.int32, device=dev))

    def one(k):
        S = sets[k % NS]
        S["run"]()
        if two_d and world > 1:
            eng.gather(S["out"], total, root=0, block=SHARD_BLOCK, out=g_out)

    for k in range(max(2, NS)):
        one(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for k in range(steps):
        one(k)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    S0 = sets[0]
    g_est, g_dlt, g_nev, g_st = (x.cpu().numpy() for x in S0["out"])  # as written by the timed region
    g_nev = g_nev.astype(np.int64)
    g_st = g_st.astype(np.uint32)
    acc = torch.tensor([float((g_st == 0).sum()), float(g_nev.sum()), float(len(g_st))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    conv, evals, count = (float(x) for x in acc.tolist())
    fl_eval = eng.functor(W.dim, W.functor)[3]
    F_t = 6.0 if not (np.isfinite(W.a) and np.isfinite(W.b)) else 0.0
    w_eval = fl_eval + F_t + 8.0
    tfl = evals * w_eval / (ms * 1e-3) / 1e12 / world  # per GPU
    res = {"workload": f"{W.name}: {W.description}; {W.algo}, {W.tol}, max_evals {W.max_evals}",
           "global_batch": int(count), "scaling": "strong" if two_d else "weak",
           "value": conv / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "converged_fraction": conv / count, "mean_nevals": evals / count, "evals_per_s": evals / (ms * 1e-3),
           "roofline": {"bound": "fp64", "achieved": tfl, "peak": peak_flops / 1e12, "unit": "TFLOP/s",
                        "frac": tfl / (peak_flops / 1e12), "flops_per_eval": w_eval, "scope": "whole batch call, per GPU"}}
    if two_d and world > 1:
        res["gather"] = "gkq_gather to rank 0 inside every timed step (20 B/integral)"
        if rank == 0:
            idx0 = torch.as_tensor(np.asarray(shard_params(W, cn, 0, world, total=total)[1]), device=dev)
            res["gathered_matches_local"] = bool(torch.equal(g_out[0][idx0], S0["out"][0]) and torch.equal(g_out[2][idx0], S0["out"][2]))
    # per-kernel profile of one call
    try:
        eng.set_profiling(True)
        flush.zero_()
        S0["run"]()
        res["roofline"]["kernels"] = [{"kernel": nm, "ms": ms_k} for nm, ms_k in eng.profile()]
        eng.set_profiling(False)
        eng.sync()
    except Exception as exc:
        res["roofline"]["kernels_error"] = repr(exc)[:120]
        eng.set_profiling(False)
    # e2e through the host entry (pinned host buffers)
    if not args.no_e2e:
        res["e2e"] = e2e_secondary(eng, W, S0["p_host"], S0["m"], world, dev, min(steps, 4))
    # CPU oracle on the same batch + parity of what the timed region wrote
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = S0["m"]
        v, cores, reps, ref = cpu_leg(W, S0["p_host"], m, seconds_target=2.0, min_reps=1, max_reps=3)
        res["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"the same {m}-integral batch, best of {reps} pass(es), all host threads"}
        res["parity_vs_cpu"] = parity(W, g_est, g_nev, g_st, ref)
    return res


def e2e_secondary(eng, W, p_host, m, world, dev, steps):
    import torch
    import torch.distributed as dist
    ph = torch.as_tensor(p_host).pin_memory()
    bufs = (torch.empty(m, dtype=torch.float64).pin_memory(), torch.empty(m, dtype=torch.float64).pin_memory(),
            torch.empty(m, dtype=torch.int32).pin_memory(), torch.empty(m, dtype=torch.int32).pin_memory())
    h2d = int(p_host.nbytes)
    if W.dim == 1:
        kw = dict(algo=ALGO[W.algo], tol=W.tol, max_evals=W.max_evals, n=m,
                  out=(bufs[0].numpy(), bufs[1].numpy(), bufs[2].numpy().view(np.uint32), bufs[3].numpy().view(np.uint32)))
        if W.points == "param0":
            kw["pts_offsets"] = np.arange(m + 1, dtype=np.int64)
            kw["pts"] = ph.numpy()[0]
            h2d += 8 * (m + 1) + 8 * m
        elif W.points:
            kw["points"] = W.points
        plan = eng.plan_batch(W.functor, W.a, W.b, ph.numpy(), **kw)
        run = plan.run
        api = "gkq_integrate_batch_host"
        status = lambda r: kw["out"][3]
    else:
        kw = dict(algo=ALGO[W.algo], tol=W.tol, max_evals=W.max_evals, n=m)
        if W.points == "param01":
            kw["pts_offsets"] = np.arange(m + 1, dtype=np.int64)
            kw["pts_xy"] = np.ascontiguousarray(p_host.T)
            h2d += 8 * (m + 1) + 16 * m
        elif W.points:
            kw["points"] = W.points

        def run():
            return eng.integrate2_batch(W.functor, W.yrange, W.a, W.b, fparams=ph.numpy(), yparams=W.yparams, **kw)
        api = "gkq_integrate2_batch_host"
        status = lambda r: r.status
    r = run()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        r = run()
    dt = (time.perf_counter() - t0) / steps
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    ce = torch.tensor([float((np.asarray(status(r)) == 0).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ce, op=dist.ReduceOp.SUM)
    return {"value": float(ce.item()) / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": int(24 * m), "ms_per_step": float(tt.item()) * 1e3, "api": api}


if __name__ == "__main__":
    main()
