"""Engine: one handle of the C-ABI library per CUDA device + the batch entry points.

torch is used here only for device memory, streams and (in distributed.py) NCCL plumbing; all
computation happens in the hand-written kernels behind include/gkquad_b200.h.
"""
from __future__ import annotations

import ctypes as C
import math
import threading

import numpy as np

from . import _ffi
from .common import BatchResult, Tolerance

ALGO_AUTO, ALGO_QAGS, ALGO_QAGP, ALGO_QAG = 0, 1, 2, 3
_engines = {}
_lock = threading.Lock()


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class BatchPlan:
    """A marshalled batch call (see Engine.plan_batch): run() crosses the C ABI and nothing else."""

    __slots__ = ("eng", "fn", "args", "device_entry", "result", "_keep")

    def __init__(self, eng, fn, args, device_entry, result, keep):
        self.eng, self.fn, self.args, self.device_entry, self.result, self._keep = eng, fn, args, device_entry, result, keep

    def run(self, stream=None):
        """Device entry: enqueue on `stream` (a raw cudaStream_t handle; default: torch's current
        stream).  Host entry: blocking.  Returns the BatchResult the plan writes into."""
        if self.device_entry:
            rc = self.fn(*self.args, self.eng._stream() if stream is None else stream)
        else:
            rc = self.fn(*self.args)
        if rc != 0:
            self.eng._check(rc)
        return self.result


class Engine:
    """Owns a gkq_handle_t (device scratch, worklists, slabs).  Not thread-safe; calls are
    ordered on the CUDA stream that is current when they are made."""

    def __init__(self, device: int = 0, stream_priority: int = 0):
        self.L = _ffi.lib()
        self.device = int(device)
        h = C.c_void_p()
        # stream_priority: CUDA priority of the streams the handle owns (gkq_create_prio)
        rc = self.L.gkq_create_prio(self.device, int(stream_priority), C.byref(h))
        if rc != 0:
            raise _ffi.GkqError(rc, self.L.gkq_last_error(None).decode())
        self.h = h
        self.nranks, self.rank = 1, 0
        self._tables = {d: _ffi.functor_table(d) for d in (0, 1, 2)}

    def close(self):
        if getattr(self, "h", None):
            self.L.gkq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- registry -------------------------------------------------------------------------
    def functor(self, dim: int, name):
        if isinstance(name, int):
            return self._tables[dim][name]
        for row in self._tables[dim]:
            if row[1] == name:
                return row
        raise KeyError(f"no {dim}-D functor named {name!r}; have {[r[1] for r in self._tables[dim]]}")

    # ---- helpers --------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise _ffi.GkqError(rc, self.L.gkq_last_error(self.h).decode())

    def _config(self, algo, tol: Tolerance, max_evals, points, workspace_capacity, dim):
        if tol.contains_nan():
            raise ValueError("Tolerance must not contain NAN.")  # integrator.rs:71
        cfg = _ffi.gkq_config()
        cfg.algo = int(algo)
        cfg.tol.kind, cfg.tol.abs_tol, cfg.tol.rel_tol = tol.kind, tol.abs_tol, tol.rel_tol
        cfg.max_evals = int(max_evals)
        cfg.workspace_capacity = int(workspace_capacity)
        flat = np.ascontiguousarray(np.asarray(points if points is not None else [], dtype=np.float64).reshape(-1))
        if np.isnan(flat).any():
            raise ValueError(f"cannot include NAN value in singular points: {points}")  # integrator.rs:86-90
        cfg.npoints = flat.size // dim
        cfg.points = flat.ctypes.data_as(C.POINTER(C.c_double)) if flat.size else None
        cfg._keepalive = flat
        return cfg

    @staticmethod
    def _ptr(x):
        if x is None:
            return None
        if isinstance(x, np.ndarray):
            return x.ctypes.data
        return x.data_ptr()

    def _stream(self):
        import torch
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- 1-D ------------------------------------------------------------------------------
    def integrate_batch(self, functor, a, b, params=None, **kw):
        """n independent integrals of `functor`; see plan_batch for the arguments.  Returns
        BatchResult (numpy arrays or torch tensors, like the inputs)."""
        return self.plan_batch(functor, a, b, params, **kw).run()

    def plan_batch(self, functor, a, b, params=None, *, algo=ALGO_AUTO, tol=None, max_evals=2000,
                   points=None, pts_offsets=None, pts=None, workspace_capacity=0, n=None, out=None,
                   abs_tol_each=None, rel_tol_each=None, max_evals_each=None):
        """Validate and marshal one batch call once; BatchPlan.run() then only crosses the C ABI
        (the counterpart of an Integrator that owns its integrand and config and is run many times).
        n independent integrals of `functor` (registry name or id).

        a, b: scalars (one shared range; `n` or a params batch defines the size) or arrays [n].
        params: None, a sequence of scalars (shared), or an array [nparams, n] (SoA).
        Arrays may be numpy (host entry: staged through pinned/pageable copies by the library) or
        torch CUDA float64 tensors (device entry: asynchronous on the current stream).
        out = (estimate, delta, nevals, status) to write into caller-owned arrays.
        abs_tol_each / rel_tol_each / max_evals_each: optional arrays [n] -- every integral's own
        tolerance values and evaluation budget (an Integrator configured anew for each integral of a
        sweep); the tolerance kind stays that of `tol`, budgets are clamped to `max_evals`."""
        fid, fname, npar, _ = self.functor(1, functor)
        tol = tol or Tolerance.default()
        on_dev = any(_is_torch(x) for x in (a, b, params) if x is not None)
        xp = None
        if on_dev:
            import torch
            xp = torch

        def arr(x, shape_tail=()):
            if x is None:
                return None
            if on_dev:
                if _is_torch(x):
                    assert x.is_cuda and x.dtype == xp.float64 and x.is_contiguous(), "need contiguous CUDA float64"
                    return x
                return xp.as_tensor(np.asarray(x, dtype=np.float64), device=f"cuda:{self.device}")
            return np.ascontiguousarray(x, dtype=np.float64)

        # ranges
        shared_range = np.isscalar(a) or (hasattr(a, "ndim") and a.ndim == 0)
        if shared_range:
            aa, bb, rs = arr([float(a)]), arr([float(b)]), 0
            if np.isnan(float(a)) or np.isnan(float(b)):
                raise ValueError("cannot create Range object from Range which contains NaN value.")
        else:
            aa, bb, rs = arr(a), arr(b), 1
            n = int(aa.shape[0])
        # params
        pp, ps = None, 0
        if npar:
            if params is None:
                raise ValueError(f"functor {fname} needs {npar} parameters")
            p = params if (_is_torch(params) or isinstance(params, np.ndarray)) else np.asarray(params, dtype=np.float64)
            if p.ndim == 1:
                assert p.shape[0] == npar
                pp, ps = arr(p), 0
            else:
                assert p.shape[0] == npar, f"params must be [nparams={npar}, n]"
                pp = arr(p)
                if n is None:
                    n = int(p.shape[1])
                assert int(p.shape[1]) == n
                ps = n
        if n is None:
            n = 1
        cfg = self._config(algo, tol, max_evals, points, workspace_capacity, 1)
        po = pv = None
        if pts_offsets is not None:
            if on_dev:
                po = pts_offsets if _is_torch(pts_offsets) else xp.as_tensor(np.asarray(pts_offsets, dtype=np.int64), device=f"cuda:{self.device}")
                pv = arr(pts)
            else:
                po = np.ascontiguousarray(pts_offsets, dtype=np.int64)
                pv = arr(pts)
        # per-integral overrides
        ov = None
        keep_ov = ()
        if abs_tol_each is not None or rel_tol_each is not None or max_evals_each is not None:
            ate, rte = arr(abs_tol_each), arr(rel_tol_each)
            mee = None
            if max_evals_each is not None:
                if on_dev:
                    mee = (max_evals_each if _is_torch(max_evals_each)
                           else xp.as_tensor(np.asarray(max_evals_each, dtype=np.int64), device=f"cuda:{self.device}"))
                    assert mee.dtype == xp.int64 and mee.is_contiguous()
                else:
                    mee = np.ascontiguousarray(max_evals_each, dtype=np.uint64)
            for x in (ate, rte, mee):
                assert x is None or int(x.shape[0]) == n, "per-integral arrays need n entries"
            ov = _ffi.gkq_overrides(self._ptr(ate), self._ptr(rte), self._ptr(mee))
            keep_ov = (ate, rte, mee, ov)
        if on_dev:
            if out is None:
                dev = f"cuda:{self.device}"
                est = xp.empty(n, dtype=xp.float64, device=dev)
                dlt = xp.empty(n, dtype=xp.float64, device=dev)
                nev = xp.empty(n, dtype=xp.int32, device=dev)
                st = xp.empty(n, dtype=xp.int32, device=dev)
            else:
                est, dlt, nev, st = out
            r = BatchResult(est, dlt, nev, st)
            if ov is not None:
                args = (self.h, C.byref(cfg), C.byref(ov), fid, n, self._ptr(aa), self._ptr(bb), rs, self._ptr(pp),
                        ps, self._ptr(po), self._ptr(pv), self._ptr(est), self._ptr(dlt), self._ptr(nev),
                        self._ptr(st))
                return BatchPlan(self, self.L.gkq_integrate_batch_ex, args, True, r,
                                 (aa, bb, pp, po, pv, cfg) + keep_ov)
            args = (self.h, C.byref(cfg), fid, n, self._ptr(aa), self._ptr(bb), rs, self._ptr(pp), ps,
                    self._ptr(po), self._ptr(pv), self._ptr(est), self._ptr(dlt), self._ptr(nev), self._ptr(st))
            return BatchPlan(self, self.L.gkq_integrate_batch, args, True, r, (aa, bb, pp, po, pv, cfg))
        if out is None:
            est, dlt = np.empty(n), np.empty(n)
            nev, st = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint32)
        else:
            est, dlt, nev, st = out
        r = BatchResult(est, dlt, nev, st)
        if ov is not None:
            args = (self.h, C.byref(cfg), C.byref(ov), fid, n, self._ptr(aa), self._ptr(bb), rs, self._ptr(pp), ps,
                    self._ptr(po), self._ptr(pv), self._ptr(est), self._ptr(dlt), self._ptr(nev), self._ptr(st))
            return BatchPlan(self, self.L.gkq_integrate_batch_host_ex, args, False, r,
                             (aa, bb, pp, po, pv, cfg) + keep_ov)
        args = (self.h, C.byref(cfg), fid, n, self._ptr(aa), self._ptr(bb), rs, self._ptr(pp), ps,
                self._ptr(po), self._ptr(pv), self._ptr(est), self._ptr(dlt), self._ptr(nev), self._ptr(st))
        return BatchPlan(self, self.L.gkq_integrate_batch_host, args, False, r, (aa, bb, pp, po, pv, cfg))

    def qk_batch(self, rule: int, functor, a, b, params=None):
        """single::qk17 ... qk57 for a batch (`rule` = number of points; host arrays in, [4, n] device tensor out)."""
        import torch
        fid, fname, npar, _ = self.functor(1, functor)
        dev = f"cuda:{self.device}"
        aa = torch.as_tensor(np.atleast_1d(np.asarray(a, dtype=np.float64)), device=dev)
        bb = torch.as_tensor(np.atleast_1d(np.asarray(b, dtype=np.float64)), device=dev)
        n = int(aa.shape[0])
        pp, ps = None, 0
        if npar:
            p = np.asarray(params, dtype=np.float64)
            pp = torch.as_tensor(np.ascontiguousarray(p), device=dev)
            ps = n if p.ndim == 2 else 0
        out = torch.empty((4, n), dtype=torch.float64, device=dev)
        self._check(self.L.gkq_qk_batch(self.h, rule, fid, n, aa.data_ptr(), bb.data_ptr(), 1,
                                        self._ptr(pp), ps, out.data_ptr(), self._stream()))
        return out

    def eval_batch(self, functor, x, params=None, mode=0):
        """y[i] = f(x[i]; params[:, i]) on the device (Integrand::apply_to_slice).  mode 1 = the sample as
        the rule kernels take it (branch-free math first; gkq_functor_eval_batch_ex)."""
        import torch
        fid, fname, npar, _ = self.functor(1, functor)
        dev = f"cuda:{self.device}"
        xx = x if _is_torch(x) else torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
        n = int(xx.shape[0])
        pp, ps = None, 0
        if npar:
            pp = params if _is_torch(params) else torch.as_tensor(
                np.ascontiguousarray(np.asarray(params, dtype=np.float64)), device=dev)
            ps = n if pp.ndim == 2 else 0
        y = torch.empty(n, dtype=torch.float64, device=dev)
        self._check(self.L.gkq_functor_eval_batch_ex(self.h, fid, n, xx.data_ptr(), self._ptr(pp), ps,
                                                     y.data_ptr(), int(mode), self._stream()))
        return y

    # ---- 2-D ------------------------------------------------------------------------------
    def integrate2_batch(self, functor2, yrange, x0, x1, fparams=None, yparams=None, *, algo=ALGO_AUTO,
                         tol=None, max_evals=100000, points=None, pts_offsets=None, pts_xy=None, n=None,
                         swap_xy=False, abs_tol_each=None, rel_tol_each=None, max_evals_each=None):
        """n nested double integrals; yrange = registry name of the y(x) functor ('const',
        'zero_to_x', 'circle'); points = list of (x, y) pairs shared by the batch, or per-integral
        CSR: pts_offsets[n+1] (counting pairs) + pts_xy[m, 2].
        swap_xy (GKQ_FLAG_SWAP_XY, the reference's DynamicX): the integrand is evaluated as
        f(inner, outer); x0/x1 and the range functor then describe the outer / inner variable and the
        points are (outer, inner) pairs.
        abs_tol_each / rel_tol_each / max_evals_each: optional arrays [n], every nest's own tolerance
        values and TOTAL evaluation budget (clamped to `max_evals`)."""
        fid, fname, npar, _ = self.functor(2, functor2)
        yid, yname, nq, _ = self.functor(0, yrange)
        tol = tol or Tolerance.default()
        on_dev = any(_is_torch(x) for x in (x0, x1, fparams, yparams) if x is not None)
        if on_dev:
            import torch

        def arr(x):
            if x is None:
                return None
            if on_dev:
                if _is_torch(x):
                    assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
                    return x
                return torch.as_tensor(np.asarray(x, dtype=np.float64), device=f"cuda:{self.device}")
            return np.ascontiguousarray(x, dtype=np.float64)

        if np.isscalar(x0):
            aa, bb, rs = arr([float(x0)]), arr([float(x1)]), 0
        else:
            aa, bb, rs = arr(x0), arr(x1), 1
            n = int(aa.shape[0])

        def par(p, cnt, what):
            nonlocal n
            if not cnt:
                return None, 0
            if p is None:
                raise ValueError(f"{what} needs {cnt} parameters")
            q = p if (_is_torch(p) or isinstance(p, np.ndarray)) else np.asarray(p, dtype=np.float64)
            if q.ndim == 1:
                assert q.shape[0] == cnt
                return arr(q), 0
            assert q.shape[0] == cnt
            if n is None:
                n = int(q.shape[1])
            assert int(q.shape[1]) == n
            return arr(q), n

        fp, fs = par(fparams, npar, fname)
        yp, ys = par(yparams, nq, yname)
        if n is None:
            n = 1
        cfg = self._config(algo, tol, max_evals, points, 0, 2)
        cfg.flags = 1 if swap_xy else 0
        po = pv = None
        if pts_offsets is not None:
            if on_dev:
                po = pts_offsets if _is_torch(pts_offsets) else torch.as_tensor(
                    np.asarray(pts_offsets, dtype=np.int64), device=f"cuda:{self.device}")
                pv = arr(pts_xy)
            else:
                po = np.ascontiguousarray(pts_offsets, dtype=np.int64)
                pv = np.ascontiguousarray(pts_xy, dtype=np.float64)
        ov = None
        keep_ov = ()
        if abs_tol_each is not None or rel_tol_each is not None or max_evals_each is not None:
            ate, rte = arr(abs_tol_each), arr(rel_tol_each)
            mee = None
            if max_evals_each is not None:
                if on_dev:
                    mee = (max_evals_each if _is_torch(max_evals_each) else torch.as_tensor(
                        np.asarray(max_evals_each, dtype=np.int64), device=f"cuda:{self.device}"))
                else:
                    mee = np.ascontiguousarray(max_evals_each, dtype=np.uint64)
            for x in (ate, rte, mee):
                assert x is None or int(x.shape[0]) == n, "per-integral arrays need n entries"
            ov = _ffi.gkq_overrides(self._ptr(ate), self._ptr(rte), self._ptr(mee))
            keep_ov = (ate, rte, mee, ov)
        ovp = C.byref(ov) if ov is not None else None
        if on_dev:
            dev = f"cuda:{self.device}"
            est = torch.empty(n, dtype=torch.float64, device=dev)
            dlt = torch.empty(n, dtype=torch.float64, device=dev)
            nev = torch.empty(n, dtype=torch.int32, device=dev)
            st = torch.empty(n, dtype=torch.int32, device=dev)
            self._check(self.L.gkq_integrate2_batch_ex(
                self.h, C.byref(cfg), ovp, fid, yid, n, self._ptr(aa), self._ptr(bb), rs, self._ptr(fp), fs,
                self._ptr(yp), ys, self._ptr(po), self._ptr(pv), self._ptr(est), self._ptr(dlt),
                self._ptr(nev), self._ptr(st), self._stream()))
            r = BatchResult(est, dlt, nev, st)
            r._keep = (aa, bb, fp, yp, po, pv, cfg) + keep_ov
            return r
        est, dlt = np.empty(n), np.empty(n)
        nev, st = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint32)
        self._check(self.L.gkq_integrate2_batch_host_ex(
            self.h, C.byref(cfg), ovp, fid, yid, n, self._ptr(aa), self._ptr(bb), rs, self._ptr(fp), fs,
            self._ptr(yp), ys, self._ptr(po), self._ptr(pv), self._ptr(est), self._ptr(dlt), self._ptr(nev),
            self._ptr(st)))
        return BatchResult(est, dlt, nev, st)

    # ---- multi-GPU (gkq_comm_* / gkq_gather: NCCL inside the library) --------------------------
    def comm_unique_id(self) -> bytes:
        """128 bytes to hand to every rank's comm_init (made on ONE rank, carried by the host)."""
        buf = C.create_string_buffer(_ffi.GKQ_COMM_ID_BYTES)
        self._check(self.L.gkq_comm_get_unique_id(buf))
        return bytes(buf.raw)

    def comm_init(self, nranks: int, rank: int, unique_id: bytes | None = None):
        """Collective: one handle per GPU joins the communicator (nranks == 1 needs no id / no NCCL)."""
        if nranks > 1:
            assert unique_id is not None and len(unique_id) == _ffi.GKQ_COMM_ID_BYTES
        self._comm_id = C.create_string_buffer(unique_id or b"", _ffi.GKQ_COMM_ID_BYTES)
        self._check(self.L.gkq_comm_init(self.h, int(nranks), int(rank), self._comm_id))
        self.nranks, self.rank = int(nranks), int(rank)

    def comm_destroy(self):
        self._check(self.L.gkq_comm_destroy(self.h))

    def shard_size(self, n_total: int, rank: int | None = None, block: int = 4096) -> int:
        return int(self.L.gkq_shard_size(int(n_total), self.nranks, self.rank if rank is None else rank, block))

    def gather(self, res, n_total: int, *, root: int = 0, block: int = 4096, out=None, stream=None,
               local_offset: int | None = None, local_count: int | None = None):
        """gkq_gather / gkq_gather_range: `res` = (estimate, delta, nevals, status) device tensors of this
        rank's shard (shard order).  Returns (estimate, delta, nevals, status) in GLOBAL order on the
        receiving rank(s) (root, or every rank when root < 0), None elsewhere.  Asynchronous on `stream`
        (default: torch's current stream)."""
        import torch
        est, dlt, nev, st = res
        receiver = root < 0 or root == self.rank
        if receiver and out is None:
            dev = f"cuda:{self.device}"
            out = (torch.empty(n_total, dtype=torch.float64, device=dev), torch.empty(n_total, dtype=torch.float64, device=dev),
                   torch.empty(n_total, dtype=torch.int32, device=dev), torch.empty(n_total, dtype=torch.int32, device=dev))
        g = out if receiver else (None, None, None, None)
        stream = self._stream() if stream is None else stream
        if local_offset is None:
            self._check(self.L.gkq_gather(self.h, int(n_total), int(block), int(root), self._ptr(est), self._ptr(dlt),
                                          self._ptr(nev), self._ptr(st), self._ptr(g[0]), self._ptr(g[1]),
                                          self._ptr(g[2]), self._ptr(g[3]), stream))
        else:
            self._check(self.L.gkq_gather_range(self.h, int(n_total), int(block), int(root), int(local_offset),
                                                int(local_count), self._ptr(est), self._ptr(dlt), self._ptr(nev),
                                                self._ptr(st), self._ptr(g[0]), self._ptr(g[1]), self._ptr(g[2]),
                                                self._ptr(g[3]), stream))
        return out if receiver else None

    # ---- accounting -------------------------------------------------------------------------
    def set_deferred_join(self, on: bool):
        """Throughput pipelining of independent batches (see gkq_set_deferred_join): results of a
        batch are complete only after join() / sync()."""
        self._check(self.L.gkq_set_deferred_join(self.h, 1 if on else 0))

    def join(self):
        self._check(self.L.gkq_join(self.h, self._stream()))

    def sync(self):
        self._check(self.L.gkq_sync(self.h, self._stream()))

    def stats(self) -> dict:
        s = _ffi.gkq_stats()
        self._check(self.L.gkq_get_stats(self.h, C.byref(s)))
        out = {}
        for k, _ in s._fields_:
            v = getattr(s, k)
            out[k] = int(v) if isinstance(v, int) else [int(x) for x in v]
        return out

    def reset_stats(self):
        self._check(self.L.gkq_reset_stats(self.h))

    def set_profiling(self, on: bool):
        self._check(self.L.gkq_set_profiling(self.h, 1 if on else 0))

    def profile(self):
        """[(kernel name, ms)] of the last batch call (needs set_profiling(True) before it)."""
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        k = self.L.gkq_get_profile(self.h, 16, names, ms)
        if k < 0:
            self._check(k)
        return [(names[i].decode(), float(ms[i])) for i in range(k)]

    def measure_fp64_peak(self, repeats: int = 5):
        fl, ms = C.c_double(), C.c_double()
        self._check(self.L.gkq_measure_fp64_peak(self.h, repeats, C.byref(fl), C.byref(ms)))
        return fl.value, ms.value


def default_engine(device: int | None = None) -> Engine:
    """Process-wide engine per device (the counterpart of an Integrator owning its algorithm)."""
    if device is None:
        try:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        except Exception:
            device = 0
    with _lock:
        if device not in _engines:
            _engines[device] = Engine(device)
        return _engines[device]
