#!/usr/bin/env python
"""bench.py -- converged integrals/s (QAGS, rel_tol 1e-10) of the batched quadrature hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one pass of the hot path over one batch of synthetic parameterised integrands
(workload S1 = BASELINE.json configs[1]: 1M smooth integrands exp(-p0 x) cos(p1 x) on [0,1],
QAGS, Tolerance::Relative(1e-10), max_evals 2000).  Weak scaling: every rank owns a block-cyclic
shard of N x batch integrals; results stay sharded, the job ends with one NCCL gather (see _main).

Prints ONE JSON line on rank 0 (see DESIGN.md section 7 for every field).
"""
from __future__ import annotations

import argparse
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
    ap.add_argument("--e2e-threads", type=int, default=2,
                    help="host threads (one handle each) that make the blocking host-entry calls of the e2e leg")
    ap.add_argument("--pipeline-depth", type=int, default=3,
                    help="handles / streams the independent batches alternate between (1 = one handle, one stream)")
    ap.add_argument("--p2p-gather", action="store_true",
                    help="N > 1: also time the gather fused into the kernels (result stores go over NVLink into "
                         "rank 0's HBM through torch's symmetric memory); opt-in because its rendezvous is collective")
    return ap.parse_args()


def shard_params(W, batch, rank, world, offset=0):
    """Block-cyclic shard of the global batch of world*batch integrals;
    every rank regenerates its own parameters from the counter-based generator (`offset` shifts
    the counters: a different batch of the same family)."""
    from gkquad_b200.workloads import block_cyclic_shard
    idx = block_cyclic_shard(batch * world, rank, world) + offset
    # the generator is indexed by the global integral number: build per-block and concatenate
    blocks = []
    B = 4096
    for s in range(0, len(idx), B):
        start = int(idx[s])
        cnt = min(B, len(idx) - s)
        blocks.append(W.make_params(start, cnt))
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


def cpu_leg(W, p, n, seconds_target=6.0, nthreads=0):
    """The reference's CPU implementation of the path = the golden-vector-pinned oracle port
    (gkquad itself needs a Rust toolchain, absent here), all host threads, on the same batch."""
    from oracle import oracle as O
    algo = {"AUTO": 0, "QAGS": 1, "QAGP": 2}[W.algo]
    kw = dict(params=p, tol=tol_to_oracle(W.tol), max_evals=W.max_evals, n=n, nthreads=nthreads)
    if W.points and W.points != "param0":
        kw["points"] = W.points
    O.integrate_batch(algo, W.functor, W.a, W.b, **{**kw, "n": min(n, 65536), "params": p[:, :min(n, 65536)]})
    reps, t0, best = 0, time.perf_counter(), 1e30
    conv = 0
    while True:
        t1 = time.perf_counter()
        r = O.integrate_batch(algo, W.functor, W.a, W.b, **kw)
        dt = time.perf_counter() - t1
        best = min(best, dt)
        conv = int((r["status"] == 0).sum())
        reps += 1
        if reps >= 3 and time.perf_counter() - t0 > seconds_target:
            break
        if reps >= 50:
            break
    return conv / best, O.num_threads() if nthreads == 0 else nthreads, reps, r


def run_reference(args, rank, world, real_stdout):
    """--impl reference: the reference's CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from gkquad_b200.workloads import WORKLOADS
    W = WORKLOADS[args.workload]
    n = args.batch
    p = W.make_params(0, n)
    from oracle import oracle as O
    algo = {"AUTO": 0, "QAGS": 1, "QAGP": 2}[W.algo]
    kw = dict(params=p, tol=tol_to_oracle(W.tol), max_evals=W.max_evals, n=n, nthreads=0)
    if W.points and W.points != "param0":
        kw["points"] = W.points
    for _ in range(max(1, args.warmup)):
        r = O.integrate_batch(algo, W.functor, W.a, W.b, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = O.integrate_batch(algo, W.functor, W.a, W.b, **kw)
    dt = (time.perf_counter() - t0) / args.steps
    conv = int((r["status"] == 0).sum())
    val = conv / dt
    cores = O.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{W.name}: {W.description}; {W.algo}, {W.tol}, max_evals {W.max_evals}",
                   "batch": n, "parallelism": f"{cores} host threads"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"the full {n}-integral batch per step (golden-vector-pinned C++ port of "
                                   "gkquad's CPU path; gkquad itself needs a Rust toolchain, absent here)"},
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


def _main(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world, real_stdout)
        return

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    import gkquad_b200 as gk
    from gkquad_b200.workloads import WORKLOADS

    W = WORKLOADS[args.workload]
    assert W.dim == 1, "bench.py times the 1-D hot path (the 2-D configs are parity-test cases)"
    n = args.batch
    algo = {"AUTO": 0, "QAGS": 1, "QAGP": 2}[W.algo]
    # throughput mode: `depth` handles on `depth` streams, batch k on lane k % depth (BatchPipeline);
    # lane 0's handle also serves the per-kernel profiling pass and the host entry (e2e)
    pipe = gk.BatchPipeline(local_rank, depth=args.pipeline_depth)
    DEPTH = pipe.depth
    eng = pipe.engines[0]
    from gkquad_b200 import distributed as D

    # 2 groups x G independent batches are cycled through, so that a step never finds its inputs or
    # outputs in the 126 MB L2 (8 x (16 MB in + 24 MB out) = 320 MB): "inputs larger than L2".
    # For N > 1 the results of a group are gathered with ONE all_gather (24 B/integral, packed) on a
    # side stream while the other group is being computed.
    G, NGROUPS = 4, 2
    NSETS = G * NGROUPS
    groups = []
    sets = []
    for gi in range(NGROUPS):
        packed = torch.zeros(G, 3 * n, dtype=torch.float64, device=dev)
        gathered = torch.empty(world, G, 3 * n, dtype=torch.float64, device=dev) if world > 1 else None
        groups.append(dict(packed=packed, gathered=gathered, done=None))
        for si in range(G):
            s = gi * G + si
            p_host, idx = shard_params(W, n, rank, world, offset=s * n * world)
            kw = dict(algo=algo, tol=W.tol, max_evals=W.max_evals, n=n)
            if W.points and W.points != "param0":
                kw["points"] = W.points
            elif W.points == "param0":
                from gkquad_b200.workloads import csr_points_from_param0
                off, pts = csr_points_from_param0(p_host)
                kw["pts_offsets"] = torch.as_tensor(off, device=dev)
                kw["pts"] = torch.as_tensor(pts, device=dev)
            # results are written straight into the packed gather buffer: est | delta | nevals,status
            row = packed[si]
            ints = row[2 * n:].view(torch.int32)
            out = (row[:n], row[n:2 * n], ints[:n], ints[n:])
            p_dev = torch.as_tensor(p_host, device=dev)
            # the call is marshalled once (an Integrator owns integrand + config and is run many times)
            plan = pipe.plan_batch(s % DEPTH, W.functor, W.a, W.b, p_dev, out=out, **kw)
            sets.append(dict(p_host=p_host, p_dev=p_dev, kw=kw, plan=plan, out=out, group=gi, lane=s % DEPTH))
    p_host = sets[0]["p_host"]
    comm = torch.cuda.Stream(device=dev) if world > 1 else None

    def step(k, gather):
        """One pass of the hot path over one batch: the rank integrates its shard and leaves the results
        in its own HBM.  With `gather` (N > 1) every G-th step also hands its group to the side stream:
        finish the group's stragglers (join) and all_gather the packed results, while the next group
        is computed."""
        s = k % NSETS
        S = sets[s]
        Gp = groups[S["group"]]
        if s % G == 0 and Gp["done"] is not None:  # the gather that last read this group must be over
            for ls in pipe.streams:
                ls.wait_event(Gp["done"])
            Gp["done"] = None
        pipe.run(S["plan"], S["lane"])
        if gather and s % G == G - 1:
            for ls in pipe.streams:
                ev = torch.cuda.Event()
                ev.record(ls)
                comm.wait_event(ev)
            with torch.cuda.stream(comm):
                for e in pipe.engines:
                    e.join()  # the gather also waits for the deferred remainders of the group
                dist.all_gather_into_tensor(Gp["gathered"], Gp["packed"])
                Gp["done"] = torch.cuda.Event()
                Gp["done"].record()

    def join():
        pipe.join()  # outstanding deferred remainders; the current stream waits for every lane
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)

    def final_gather(k_last):
        """north_star's 'final NCCL gather of results and status': the last batch, to every rank."""
        S = sets[k_last % NSETS]
        row = groups[S["group"]]["packed"][(k_last % NSETS) % G]
        dist.all_gather_into_tensor(last_gathered, row)

    # throughput pipelining of independent batches: the latency-bound remainders of several batches
    # are finished by one launch (deferred join); every step is complete before the end event
    last_gathered = torch.empty(world, 3 * n, dtype=torch.float64, device=dev) if world > 1 else None

    steps = args.steps
    if world > 1 and steps % G:
        steps += G - steps % G  # whole groups (the per-step-gather leg gathers group-wise)
    # nvidia-smi is started BEFORE the warm-up: while it attaches to the driver, CUDA launches of this
    # process stall for milliseconds, which must not fall into the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    # (at least two deferred-join cycles of 8 batches per lane: every scratch buffer has reached its
    # steady-state size before the timed region)
    for k in range(max(18 * DEPTH, 2 * NSETS, args.warmup, 3)):
        step(k, world > 1)
    join()
    if world > 1:
        final_gather(0)
    torch.cuda.synchronize()

    # ---- timed region: EXACTLY K steps between barrier + synchronize, device-timed -------------
    # The path shards with no data-path collective: each rank's results stay in its HBM; the job ends
    # with the final gather of the last batch (inside the region).
    pipe.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    pipe.fork()  # the lanes start behind the start event
    t_region0 = time.perf_counter()
    for k in range(steps):
        step(k, False)
    t_enq = (time.perf_counter() - t_region0) / steps
    join()  # ... and the end event behind every lane
    if world > 1:
        final_gather(steps - 1)
    e1.record()
    torch.cuda.synchronize()
    t_region1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    launches = pipe.kernel_launches()  # our kernels only (NCCL's are not counted)
    total_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    conv = sum(float((S["out"][3] == 0).sum().item()) for S in sets) / NSETS  # mean per step
    conv_local = torch.tensor([conv], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(conv_local, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    converged = float(conv_local.item())
    ms_per_step = total_ms / steps
    value = converged / (ms_per_step * 1e-3)
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None

    # ---- N > 1 only: the same K steps with EVERY batch gathered to every rank (group-wise, on a side
    # stream).  Reported next to `value`: at this per-GPU rate a full gather needs 24 B x 6e9/s x (N-1)
    # of NVLink ingress per GPU, which passes the 900 GB/s of one GPU between N = 4 and N = 8.
    per_step_gather = None
    if world > 1:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        g0.record()
        pipe.fork()
        for k in range(steps):
            step(k, True)
        join()
        g1.record()
        torch.cuda.synchronize()
        dist.barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        ms_g = float(tg.item()) / steps
        per_step_gather = {"value": converged / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g,
                           "gathered_bytes_per_step_per_gpu": int(24 * n * world),
                           "note": "every batch all_gathered (24 B/integral) inside the timed region, group-wise on a "
                                   "side stream; bound by NVLink ingress, not by the integration kernels"}
    torch.cuda.synchronize()
    for e in pipe.engines:
        e.set_deferred_join(False)
    p_dev, kw = sets[0]["p_dev"], sets[0]["kw"]
    est, dlt, nev, st = sets[0]["out"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # L2 flush for the profiling pass

    # ---- N > 1 only: the gather fused into the kernels -- every rank's kernels store their results
    # straight into RANK 0's buffer over NVLink (peer mapping from torch's symmetric memory), so the
    # transfer overlaps the math store by store and no collective kernel runs at all.
    p2p_gather = None
    if world > 1 and args.p2p_gather:
        try:
            import torch.distributed._symmetric_memory as symm
            sbuf = symm.empty((NSETS, world, 3 * n), dtype=torch.float64, device=dev)
            hdl = symm.rendezvous(sbuf, dist.group.WORLD)
            root = hdl.get_buffer(0, (NSETS, world, 3 * n), torch.float64)
            pplans = []
            for s_i, S in enumerate(sets):
                row = root[s_i, rank]
                ints = row[2 * n:].view(torch.int32)
                pplans.append(eng.plan_batch(W.functor, W.a, W.b, S["p_dev"],
                                             out=(row[:n], row[n:2 * n], ints[:n], ints[n:]), **S["kw"]))
            eng.set_deferred_join(True)
            for k in range(2 * NSETS):
                pplans[k % NSETS].run()
            eng.join()
            torch.cuda.synchronize()
            hdl.barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            q0.record()
            for k in range(steps):
                pplans[k % NSETS].run()
            eng.join()
            q1.record()
            torch.cuda.synchronize()
            hdl.barrier()  # every rank's stores have landed on rank 0
            dist.barrier()
            tq = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device=dev)
            dist.all_reduce(tq, op=dist.ReduceOp.MAX)
            ms_q = float(tq.item()) / steps
            landed = None
            if rank == 0:  # what arrived on rank 0 is what the ranks computed
                ints0 = sbuf[0, :, 2 * n:].reshape(world, -1).view(torch.int32)
                landed = float((ints0[:, n:2 * n] == 0).sum().item())
            eng.set_deferred_join(False)
            p2p_gather = {"value": converged / (ms_q * 1e-3), "unit": UNIT, "ms_per_step": ms_q,
                          "converged_seen_on_rank0_last_batch": landed,
                          "note": "every batch gathered to rank 0 by the integration kernels themselves: their result "
                                  "stores go over NVLink into rank 0's HBM (symmetric-memory peer mapping), no "
                                  "collective kernel; bound by rank 0's NVLink ingress for N > 6"}
        except Exception as exc:  # symmetric memory not available on this box
            p2p_gather = {"unavailable": repr(exc)[:200]}
            eng.set_deferred_join(False)

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
    nev_h = nev.cpu().numpy().astype(np.int64)
    fl_eval = eng.functor(1, W.functor)[3]
    F_t = 6.0 if not (np.isfinite(W.a) and np.isfinite(W.b)) else 0.0
    w_eval = fl_eval + F_t + 8.0  # W_i = nevals_i * (F_f + F_t + 8)
    n25 = stats["last_survivors_qk25"]
    cl, c2 = stats["last_chain_loop"], stats["last_chain_step2"]
    evals_by_kernel = {}
    if W.algo == "QAGS":
        rest = int(nev_h.sum()) - 17 * n - 25 * n25 - 50 * (cl[0] + cl[1])
        tail_share = [c2[k] / max(1, c2[0] + c2[1]) for k in range(2)]
        evals_by_kernel = {"k_qags_init<qk17>": 17 * n, "k_qags_init<qk25>": 25 * n25,
                           "k_qags_first[0]": 50 * cl[0], "k_qags_first[1]": 50 * cl[1],
                           "k_adaptive<QAGS>[0]": int(rest * tail_share[0]),
                           "k_adaptive<QAGS>[1]": int(rest * tail_share[1])}
    else:
        evals_by_kernel = {"k_adaptive<QAGP>": int(nev_h.sum())}
    kernels = []
    for name, v in prof_acc.items():
        ms = float(np.mean(v))
        fl = evals_by_kernel.get(name, 0) * w_eval
        kernels.append({"kernel": name, "ms": ms, "evals": evals_by_kernel.get(name, 0),
                        "tflops": fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0})
    peak_flops, peak_ms = eng.measure_fp64_peak(5)
    dom = max(kernels, key=lambda k: k["ms"]) if kernels else None
    whole_tflops = float(nev_h.sum()) * w_eval / (ms_per_step * 1e-3) / 1e12
    roofline = None
    if dom:
        roofline = {
            "bound": "fp64", "kernel": dom["kernel"], "achieved": dom["tflops"], "peak": peak_flops / 1e12,
            "unit": "TFLOP/s", "frac": dom["tflops"] / (peak_flops / 1e12),
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of that kernel, from the ncu
            # --set full capture of this workload (profiles/ncu_S1_kernels_r01d_raw.csv: 37.9 MB read +
            # 21.1 MB written, against 16.8 MB of parameters in and ~21 MB of results / parked state /
            # worklists out: partial-sector result stores make L2 fetch the rest of their lines); bytes
            "traffic": 59.0e6 if (args.workload == "S1" and n == 1 << 20 and dom["kernel"] == "k_qags_init<qk17>") else None,
            "peak_source": "measured in this run: dependent-chain DFMA micro-kernel over all SMs "
                           "(MEASURED_PEAKS.json carries no FP64 figure; BASELINE.md section 3)",
            "flops_per_eval": w_eval, "kernels": kernels,
            "whole_step": {"achieved": whole_tflops, "frac": whole_tflops / (peak_flops / 1e12)},
            "hbm": {"algorithmic_bytes_per_step": int(n * (8 * p_host.shape[0] + 24)),
                    "note": "HBM is two orders of magnitude from binding (SURVEY.md section 8d)",
                    # north_star's "worklist compaction pass" is not a pass here: survivors are appended to
                    # the worklists by warp-aggregated atomics inside the rule kernels (4 B written per
                    # survivor, 4 B read back by the consumer kernel)
                    "compaction": {"survivor_ids_bytes_per_step": int(8 * (n25 + cl[0] + cl[1])),
                                   "survivors_qk25": int(n25), "survivors_loop": int(cl[0] + cl[1]),
                                   "note": "fused into k_qags_init / k_qags_first (ballot + popc + one atomic per "
                                           "warp per list); removing the appends altogether shortens "
                                           "k_qags_init<qk17> by 4 % (profiles/S1_r01.md)"}},
        }

    # ---- e2e: the public host entry (gkq_integrate_batch_host) with pinned host buffers ----------
    e2e = None
    if not args.no_e2e:
        ph = torch.as_tensor(p_host).pin_memory()
        oe = torch.empty(n, dtype=torch.float64).pin_memory()
        od = torch.empty(n, dtype=torch.float64).pin_memory()
        on = torch.empty(n, dtype=torch.int32).pin_memory()
        os_ = torch.empty(n, dtype=torch.int32).pin_memory()
        out = (oe.numpy(), od.numpy(), on.numpy().view(np.uint32), os_.numpy().view(np.uint32))
        kwh = dict(algo=algo, tol=W.tol, max_evals=W.max_evals, n=n, out=out)
        if W.points and W.points != "param0":
            kwh["points"] = W.points
        if W.points != "param0":
            # `--e2e-threads` host threads, each with its own handle and its own pinned result
            # buffers, make the blocking call in turn (ctypes releases the GIL): the H2D of one call
            # overlaps the D2H of the other.  Every call still copies its step's inputs in and its
            # results out inside the timed region.
            nth = max(1, min(args.e2e_threads, DEPTH))
            hplans, houts = [], []
            for t in range(nth):
                if t == 0:
                    o_t = out
                else:
                    bufs = (torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
                            torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory())
                    houts.append(bufs)
                    o_t = (bufs[0].numpy(), bufs[1].numpy(), bufs[2].numpy().view(np.uint32), bufs[3].numpy().view(np.uint32))
                hplans.append(pipe.engines[t].plan_batch(W.functor, W.a, W.b, ph.numpy(), **dict(kwh, out=o_t)))
            for hp_ in hplans:
                for _ in range(3):
                    hp_.run()
            hplan = hplans[0]
            if world > 1:
                dist.barrier()
            e2e_steps = args.steps - args.steps % nth if args.steps >= nth else nth

            def host_worker(t):
                for _ in range(e2e_steps // nth):
                    hplans[t].run()
            workers = [threading.Thread(target=host_worker, args=(t,)) for t in range(1, nth)]
            t0 = time.perf_counter()
            for w_ in workers:
                w_.start()
            host_worker(0)
            for w_ in workers:
                w_.join()
            dt = (time.perf_counter() - t0) / e2e_steps
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            conv_e = float((out[3] == 0).sum())
            ce = torch.tensor([conv_e], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ce, op=dist.ReduceOp.SUM)
            e2e = {"value": float(ce.item()) / float(tt.item()), "unit": UNIT,
                   "h2d_bytes_per_step": int(p_host.nbytes), "d2h_bytes_per_step": int(24 * n),
                   "ms_per_step": float(tt.item()) * 1e3,
                   "api": "gkq_integrate_batch_host (pinned host buffers, chunked H2D/compute/D2H pipeline)"
                          + (f", called from {nth} host threads with one handle each" if nth > 1 else "")}

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, reps, ref = cpu_leg(W, p_host, n)
        g_est = est.cpu().numpy()
        g_st = st.cpu().numpy().astype(np.uint32)
        tol_abs = W.tol.to_abs(np.abs(ref["estimate"]))
        okrows = ref["status"] == 0
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the same {n}-integral batch, best of {reps} passes, all host threads "
                         "(golden-vector-pinned C++ port of gkquad's CPU path; no Rust toolchain here)",
               "parity_vs_cpu": {"status_mismatch": int((g_st != ref["status"]).sum()),
                                 "nevals_mismatch": int((nev_h != ref["nevals"]).sum()),
                                 "out_of_tolerance": int((np.abs(g_est - ref["estimate"])[okrows] > tol_abs[okrows]).sum())}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{W.name}: {W.description}; {W.algo}, {W.tol}, max_evals {W.max_evals}",
                       "batch_per_gpu": n, "global_batch": n * world,
                       "parallelism": f"shard{world}" + ("+final_nccl_all_gather" if world > 1 else ""),
                       "l2": "8 independent batches cycled (320 MB of inputs+outputs > 126 MB L2); no flush needed",
                       "pipeline": f"BatchPipeline depth {DEPTH}: independent batches alternate between {DEPTH} handles "
                                   f"on {DEPTH} streams (the single-wave survivor kernels of one batch fill the "
                                   "ramps and tails of the other); deferred join: the stragglers (integrals "
                                   "needing > 1 bisection step) of up to 8 steps of a lane are finished by one "
                                   "launch of the persistent kernel; all K steps complete inside the timed region",
                       "gather": ("no data-path collective: every rank keeps its shard's results in its own HBM; "
                                  "one final all_gather of the last batch (24 B/integral) inside the timed region; "
                                  "`per_step_gather` is the same run with every batch gathered") if world > 1
                       else "none (N=1)",
                       "converged_fraction": converged / (n * world),
                       "mean_nevals": float(nev_h.mean())},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": t_enq * 1e3,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        if per_step_gather is not None:
            line["per_step_gather"] = per_step_gather
        if p2p_gather is not None:
            line["p2p_gather"] = p2p_gather
        emit(real_stdout, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
