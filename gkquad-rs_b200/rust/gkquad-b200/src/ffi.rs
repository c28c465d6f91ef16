//! Raw bindings of include/gkquad_b200.h (ABI version 1).  UNCOMPILED in the build image.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

#[repr(C)]
pub struct gkq_handle_s {
    _private: [u8; 0],
}
pub type gkq_handle_t = *mut gkq_handle_s;

pub const GKQ_ALGO_AUTO: i32 = 0;
pub const GKQ_ALGO_QAGS: i32 = 1;
pub const GKQ_ALGO_QAGP: i32 = 2;
pub const GKQ_ALGO_QAG: i32 = 3;

pub const GKQ_TOL_ABSOLUTE: i32 = 0;
pub const GKQ_TOL_RELATIVE: i32 = 1;
pub const GKQ_TOL_ABS_OR_REL: i32 = 2;
pub const GKQ_TOL_ABS_AND_REL: i32 = 3;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct gkq_tolerance {
    pub kind: i32,
    pub _pad: i32,
    pub abs_tol: c_double,
    pub rel_tol: c_double,
}

/// Per-integral tolerance values / budgets (null = the batch value); include/gkquad_b200.h.
#[repr(C)]
pub struct gkq_overrides {
    pub abs_tol: *const c_double,
    pub rel_tol: *const c_double,
    pub max_evals: *const u64,
}

pub const GKQ_FLAG_SWAP_XY: i32 = 1;

#[repr(C)]
pub struct gkq_config {
    pub algo: i32,
    pub flags: i32,
    pub tol: gkq_tolerance,
    pub max_evals: u64,
    pub workspace_capacity: u64,
    pub points: *const c_double,
    pub npoints: i64,
}

#[repr(C)]
#[derive(Default)]
pub struct gkq_stats {
    pub kernel_launches: u64,
    pub batches: u64,
    pub integrals: u64,
    pub scratch_bytes: u64,
    pub last_survivors_qk25: u64,
    pub last_survivors_loop: u64,
    pub last_survivors_step2: u64,
    pub last_chain_loop: [u64; 2],
    pub last_chain_step2: [u64; 2],
}

extern "C" {
    pub fn gkq_abi_version() -> c_int;
    pub fn gkq_config_default(dim: c_int, cfg: *mut gkq_config) -> c_int;
    pub fn gkq_create(device: c_int, out: *mut gkq_handle_t) -> c_int;
    pub fn gkq_create_prio(device: c_int, stream_priority: c_int, out: *mut gkq_handle_t) -> c_int;
    pub fn gkq_destroy(h: gkq_handle_t) -> c_int;
    pub fn gkq_last_error(h: gkq_handle_t) -> *const c_char;
    pub fn gkq_functor_count(dim: c_int) -> c_int;
    pub fn gkq_functor_info(dim: c_int, id: c_int, name: *mut *const c_char, nparams: *mut c_int,
                            flops_per_eval: *mut c_double) -> c_int;
    pub fn gkq_functor_lookup(dim: c_int, name: *const c_char) -> c_int;
    pub fn gkq_integrate_batch(h: gkq_handle_t, cfg: *const gkq_config, functor_id: c_int, n: i64,
                               a: *const c_double, b: *const c_double, range_stride: i64,
                               params: *const c_double, param_stride: i64,
                               pts_offsets: *const i64, pts: *const c_double,
                               out_estimate: *mut c_double, out_delta: *mut c_double,
                               out_nevals: *mut u32, out_status: *mut u32, stream: *mut c_void) -> c_int;
    pub fn gkq_integrate_batch_host(h: gkq_handle_t, cfg: *const gkq_config, functor_id: c_int, n: i64,
                                    a: *const c_double, b: *const c_double, range_stride: i64,
                                    params: *const c_double, param_stride: i64,
                                    pts_offsets: *const i64, pts: *const c_double,
                                    out_estimate: *mut c_double, out_delta: *mut c_double,
                                    out_nevals: *mut u32, out_status: *mut u32) -> c_int;
    pub fn gkq_integrate2_batch(h: gkq_handle_t, cfg: *const gkq_config, functor2_id: c_int,
                                yrange_id: c_int, n: i64, x0: *const c_double, x1: *const c_double,
                                range_stride: i64, fparams: *const c_double, fparam_stride: i64,
                                yparams: *const c_double, yparam_stride: i64,
                                pts_offsets: *const i64, pts_xy: *const c_double,
                                out_estimate: *mut c_double, out_delta: *mut c_double,
                                out_nevals: *mut u32, out_status: *mut u32, stream: *mut c_void) -> c_int;
    pub fn gkq_integrate2_batch_host(h: gkq_handle_t, cfg: *const gkq_config, functor2_id: c_int,
                                     yrange_id: c_int, n: i64, x0: *const c_double, x1: *const c_double,
                                     range_stride: i64, fparams: *const c_double, fparam_stride: i64,
                                     yparams: *const c_double, yparam_stride: i64,
                                     pts_offsets: *const i64, pts_xy: *const c_double,
                                     out_estimate: *mut c_double, out_delta: *mut c_double,
                                     out_nevals: *mut u32, out_status: *mut u32) -> c_int;
    pub fn gkq_integrate_batch_ex(h: gkq_handle_t, cfg: *const gkq_config, ov: *const gkq_overrides,
        functor_id: c_int, n: i64, a: *const f64, b: *const f64, range_stride: i64, params: *const f64,
        param_stride: i64, pts_offsets: *const i64, pts: *const f64, out_estimate: *mut f64,
        out_delta: *mut f64, out_nevals: *mut u32, out_status: *mut u32, stream: *mut c_void) -> c_int;
    pub fn gkq_integrate_batch_host_ex(h: gkq_handle_t, cfg: *const gkq_config, ov: *const gkq_overrides,
        functor_id: c_int, n: i64, a: *const f64, b: *const f64, range_stride: i64, params: *const f64,
        param_stride: i64, pts_offsets: *const i64, pts: *const f64, out_estimate: *mut f64,
        out_delta: *mut f64, out_nevals: *mut u32, out_status: *mut u32) -> c_int;
    pub fn gkq_integrate2_batch_ex(h: gkq_handle_t, cfg: *const gkq_config, ov: *const gkq_overrides,
        functor2_id: c_int, yrange_id: c_int, n: i64, x0: *const f64, x1: *const f64, range_stride: i64,
        fparams: *const f64, fparam_stride: i64, yparams: *const f64, yparam_stride: i64,
        pts_offsets: *const i64, pts_xy: *const f64, out_estimate: *mut f64, out_delta: *mut f64,
        out_nevals: *mut u32, out_status: *mut u32, stream: *mut c_void) -> c_int;
    pub fn gkq_integrate2_batch_host_ex(h: gkq_handle_t, cfg: *const gkq_config, ov: *const gkq_overrides,
        functor2_id: c_int, yrange_id: c_int, n: i64, x0: *const f64, x1: *const f64, range_stride: i64,
        fparams: *const f64, fparam_stride: i64, yparams: *const f64, yparam_stride: i64,
        pts_offsets: *const i64, pts_xy: *const f64, out_estimate: *mut f64, out_delta: *mut f64,
        out_nevals: *mut u32, out_status: *mut u32) -> c_int;
    pub fn gkq_qk_batch(h: gkq_handle_t, rule: c_int, functor_id: c_int, n: i64, a: *const c_double,
                        b: *const c_double, range_stride: i64, params: *const c_double,
                        param_stride: i64, out4: *mut c_double, stream: *mut c_void) -> c_int;
    pub fn gkq_functor_eval_batch(h: gkq_handle_t, functor_id: c_int, n: i64, x: *const c_double,
                                  params: *const c_double, param_stride: i64, y: *mut c_double,
                                  stream: *mut c_void) -> c_int;
    pub fn gkq_functor_eval_batch_ex(h: gkq_handle_t, functor_id: c_int, n: i64, x: *const c_double,
                                     params: *const c_double, param_stride: i64, y: *mut c_double,
                                     mode: c_int, stream: *mut c_void) -> c_int;
    pub fn gkq_set_deferred_join(h: gkq_handle_t, enable: c_int) -> c_int;
    pub fn gkq_join(h: gkq_handle_t, stream: *mut c_void) -> c_int;
    pub fn gkq_sync(h: gkq_handle_t, stream: *mut c_void) -> c_int;
    pub fn gkq_get_stats(h: gkq_handle_t, out: *mut gkq_stats) -> c_int;
    pub fn gkq_reset_stats(h: gkq_handle_t) -> c_int;
    pub fn gkq_set_profiling(h: gkq_handle_t, enable: c_int) -> c_int;
    pub fn gkq_get_profile(h: gkq_handle_t, max_entries: c_int, names: *mut *const c_char,
                           ms: *mut c_float) -> c_int;
    pub fn gkq_measure_fp64_peak(h: gkq_handle_t, repeats: c_int, flops_per_s: *mut c_double,
                                 ms: *mut c_double) -> c_int;
    // multi-GPU: static block-cyclic shard + one gather of results (NCCL inside the library)
    pub fn gkq_comm_get_unique_id(id_out: *mut c_void) -> c_int; // [GKQ_COMM_ID_BYTES]
    pub fn gkq_comm_init(h: gkq_handle_t, nranks: c_int, rank: c_int, unique_id: *const c_void) -> c_int;
    pub fn gkq_comm_destroy(h: gkq_handle_t) -> c_int;
    pub fn gkq_comm_info(h: gkq_handle_t, nranks: *mut c_int, rank: *mut c_int) -> c_int;
    pub fn gkq_shard_size(n_total: i64, nranks: c_int, rank: c_int, block: i64) -> i64;
    pub fn gkq_shard_index(n_total: i64, nranks: c_int, rank: c_int, block: i64, j: i64) -> i64;
    pub fn gkq_gather(h: gkq_handle_t, n_total: i64, block: i64, root: c_int, est: *const f64, dlt: *const f64,
                      nev: *const u32, st: *const u32, g_est: *mut f64, g_dlt: *mut f64, g_nev: *mut u32,
                      g_st: *mut u32, stream: *mut c_void) -> c_int;
    pub fn gkq_gather_range(h: gkq_handle_t, n_total: i64, block: i64, root: c_int, local_offset: i64,
                            local_count: i64, est: *const f64, dlt: *const f64, nev: *const u32, st: *const u32,
                            g_est: *mut f64, g_dlt: *mut f64, g_nev: *mut u32, g_st: *mut u32,
                            stream: *mut c_void) -> c_int;
}
pub const GKQ_COMM_ID_BYTES: usize = 128;
pub const GKQ_FLAG_SWAP_XY: i32 = 1;
pub const GKQ_FLAG_VALIDATE_INPUTS: i32 = 2;
