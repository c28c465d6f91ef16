//! gkquad-b200: gkquad's `Algorithm` / `Algorithm2` plug-in seam implemented on the B200 engine.
//!
//! UNCOMPILED in the build image (no Rust toolchain); mirrors what the compiled C++ / Python hosts
//! do.  The seam: `trait Algorithm<F: Integrand + ?Sized>` (gkquad/src/single/algorithm/mod.rs:16-23)
//! is generic over the integrand, so a type implementing `Algorithm<DeviceIntegrand>` is a legal
//! argument of `Integrator::with_algorithm` / `.algorithm(..)` (single/integrator.rs:50-66):
//!
//! ```ignore
//! use gkquad::single::Integrator;
//! use gkquad_b200::{DeviceIntegrand, QagsB200};
//! let f = DeviceIntegrand::new("expcos_p", vec![p0, p1]);        // SoA parameter batch
//! let mut it = Integrator::with_algorithm(f, QagsB200::new(0)?)  // device 0
//!     .tolerance(gkquad::Tolerance::Relative(1e-10));
//! let first = it.run(0.0..1.0);                                   // IntegrationResult of integral 0
//! let all = it.get_algorithm().last_batch();                      // every integral of the batch
//! ```
pub mod ffi;

use gkquad::single::algorithm::Algorithm;
use gkquad::single::{Integrand, IntegrationConfig, Range};
use gkquad::{IntegrationResult, RuntimeError, Solution, Tolerance};
use std::ffi::CString;

/// A functor of the device registry plus a batch of parameters (`params[k][i]`).  It implements
/// `Integrand` only to satisfy the trait bound; the CPU never evaluates it.
pub struct DeviceIntegrand {
    pub functor: String,
    pub params: Vec<Vec<f64>>,
}
impl DeviceIntegrand {
    pub fn new(functor: &str, params: Vec<Vec<f64>>) -> Self {
        Self { functor: functor.to_string(), params }
    }
    pub fn batch(&self) -> usize {
        self.params.first().map(|r| r.len()).unwrap_or(1)
    }
}
impl Integrand for DeviceIntegrand {
    fn apply(&mut self, _x: f64) -> f64 {
        panic!("DeviceIntegrand is evaluated on the GPU; there is no CPU fallback")
    }
}

/// Structure-of-arrays results of the last batch.
#[derive(Default, Clone)]
pub struct BatchResult {
    pub estimate: Vec<f64>,
    pub delta: Vec<f64>,
    pub nevals: Vec<u32>,
    pub status: Vec<u32>,
}
impl BatchResult {
    pub fn get(&self, i: usize) -> IntegrationResult {
        let value = Solution { estimate: self.estimate[i], delta: self.delta[i], nevals: self.nevals[i] as usize };
        match self.status[i] {
            0 => IntegrationResult::new(value),
            1 => IntegrationResult::with_error(value, RuntimeError::InsufficientIteration),
            2 => IntegrationResult::with_error(value, RuntimeError::RoundoffError),
            3 => IntegrationResult::with_error(value, RuntimeError::SubrangeTooSmall),
            4 => IntegrationResult::with_error(value, RuntimeError::Divergent),
            _ => IntegrationResult::with_error(value, RuntimeError::NanValueEncountered),
        }
    }
}

fn tol_to_ffi(t: &Tolerance) -> ffi::gkq_tolerance {
    let (kind, a, r) = match *t {
        Tolerance::Absolute(x) => (ffi::GKQ_TOL_ABSOLUTE, x, 0.0),
        Tolerance::Relative(y) => (ffi::GKQ_TOL_RELATIVE, 0.0, y),
        Tolerance::AbsOrRel(x, y) => (ffi::GKQ_TOL_ABS_OR_REL, x, y),
        Tolerance::AbsAndRel(x, y) => (ffi::GKQ_TOL_ABS_AND_REL, x, y),
    };
    ffi::gkq_tolerance { kind, _pad: 0, abs_tol: a, rel_tol: r }
}

/// Owns a `gkq_handle_t` (device scratch): the counterpart of the `WorkSpace` a `QAGS<'a>` owns.
pub struct EngineB200 {
    h: ffi::gkq_handle_t,
    algo: i32,
    last: BatchResult,
}
impl EngineB200 {
    fn with_algo(device: i32, algo: i32) -> Result<Self, String> {
        Self::with_algo_prio(device, algo, 0)
    }
    /// `stream_priority`: CUDA priority of the streams the handle owns (0 = default, negative = scheduled
    /// first; `gkq_create_prio`) -- the express lane of a pipeline of engines.
    fn with_algo_prio(device: i32, algo: i32, stream_priority: i32) -> Result<Self, String> {
        let mut h: ffi::gkq_handle_t = std::ptr::null_mut();
        let rc = unsafe { ffi::gkq_create_prio(device, stream_priority, &mut h) };
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::gkq_last_error(std::ptr::null_mut())) };
            return Err(msg.to_string_lossy().into_owned());
        }
        Ok(Self { h, algo, last: BatchResult::default() })
    }
    /// Every integral of the last `integrate` call.
    pub fn last_batch(&self) -> &BatchResult {
        &self.last
    }
    fn run(&mut self, f: &DeviceIntegrand, range: &Range, config: &IntegrationConfig) -> IntegrationResult {
        let n = f.batch();
        let name = CString::new(f.functor.as_str()).unwrap();
        let fid = unsafe { ffi::gkq_functor_lookup(1, name.as_ptr()) };
        assert!(fid >= 0, "unknown device functor {}", f.functor);
        let flat: Vec<f64> = f.params.iter().flat_map(|r| r.iter().cloned()).collect();
        let cfg = ffi::gkq_config {
            algo: self.algo,
            flags: 0,
            tol: tol_to_ffi(&config.tolerance),
            max_evals: config.max_evals as u64,
            workspace_capacity: 0,
            points: if config.points.is_empty() { std::ptr::null() } else { config.points.as_ptr() },
            npoints: config.points.len() as i64,
        };
        self.last = BatchResult { estimate: vec![0.0; n], delta: vec![0.0; n], nevals: vec![0; n], status: vec![0; n] };
        let rc = unsafe {
            ffi::gkq_integrate_batch_host(
                self.h, &cfg, fid, n as i64, &range.begin, &range.end, 0,
                if flat.is_empty() { std::ptr::null() } else { flat.as_ptr() },
                if flat.is_empty() { 0 } else { n as i64 },
                std::ptr::null(), std::ptr::null(),
                self.last.estimate.as_mut_ptr(), self.last.delta.as_mut_ptr(),
                self.last.nevals.as_mut_ptr(), self.last.status.as_mut_ptr())
        };
        assert_eq!(rc, 0, "gkq_integrate_batch_host failed");
        self.last.get(0)
    }
}
impl Drop for EngineB200 {
    fn drop(&mut self) {
        unsafe { ffi::gkq_destroy(self.h) };
    }
}

macro_rules! algorithm {
    ($name:ident, $algo:expr, $doc:expr) => {
        #[doc = $doc]
        pub struct $name(EngineB200);
        impl $name {
            pub fn new(device: i32) -> Result<Self, String> {
                Ok(Self(EngineB200::with_algo(device, $algo)?))
            }
            pub fn last_batch(&self) -> &BatchResult {
                self.0.last_batch()
            }
        }
        impl Algorithm<DeviceIntegrand> for $name {
            fn integrate(&mut self, f: &mut DeviceIntegrand, range: &Range, config: &IntegrationConfig) -> IntegrationResult {
                self.0.run(f, range, config)
            }
        }
    };
}
algorithm!(QagsB200, ffi::GKQ_ALGO_QAGS, "replaces gkquad::single::algorithm::QAGS (qags.rs:19-63)");
algorithm!(QagpB200, ffi::GKQ_ALGO_QAGP, "replaces gkquad::single::algorithm::QAGP (qagp.rs:20-64)");
algorithm!(AutoB200, ffi::GKQ_ALGO_AUTO, "replaces gkquad::single::algorithm::AUTO (auto.rs:7-35)");
algorithm!(QagB200, ffi::GKQ_ALGO_QAG, "replaces the deprecated gkquad::single::algorithm::QAG (qag.rs:16-60)");

// ---- 2-D: the Algorithm2 seam (gkquad/src/double/algorithm/mod.rs:7-10) -----------------------------
use gkquad::double::algorithm::Algorithm2;
use gkquad::double::range::Rectangle;
use gkquad::double::{Integrand2, IntegrationConfig2};

/// A 2-D functor of the device registry, a y-range functor (`"const"` for a Rectangle, `"zero_to_x"`,
/// `"circle"`: the DynamicY closures of double/range.rs:77-112 cannot cross the FFI) and SoA parameter
/// batches for both.  `Integrand2` is implemented only to satisfy the trait bound.
pub struct DeviceIntegrand2 {
    pub functor: String,
    pub fparams: Vec<Vec<f64>>,
    pub yrange: String,
    pub yparams: Vec<Vec<f64>>,
    /// per-integral singular points in CSR form (offsets count (x, y) pairs); empty = config.points
    pub pts_offsets: Vec<i64>,
    pub pts_xy: Vec<f64>,
}
impl Integrand2 for DeviceIntegrand2 {
    fn apply(&mut self, _x: (f64, f64)) -> f64 {
        panic!("DeviceIntegrand2 is evaluated on the GPU; there is no CPU fallback")
    }
}
impl DeviceIntegrand2 {
    pub fn batch(&self) -> usize {
        self.fparams.first().or(self.yparams.first()).map(|r| r.len()).unwrap_or(1)
    }
}

impl EngineB200 {
    /// `gkq_integrate2_batch_host`: every nest of the batch over the x-range `xr`; the y-range comes
    /// from the integrand's y-range functor (a Rectangle passes its y-range as the two parameters of
    /// `"const"`, double/algorithm/qags.rs:18-27).
    fn run2(&mut self, f: &DeviceIntegrand2, xr: &Range, yconst: Option<&Range>, config: &IntegrationConfig2) -> IntegrationResult {
        let n = f.batch();
        let name = CString::new(f.functor.as_str()).unwrap();
        let yname = CString::new(if yconst.is_some() { "const" } else { f.yrange.as_str() }).unwrap();
        let fid = unsafe { ffi::gkq_functor_lookup(2, name.as_ptr()) };
        let yid = unsafe { ffi::gkq_functor_lookup(0, yname.as_ptr()) };
        assert!(fid >= 0 && yid >= 0, "unknown device functor {} / {}", f.functor, f.yrange);
        let fflat: Vec<f64> = f.fparams.iter().flat_map(|r| r.iter().cloned()).collect();
        let (yflat, ystride): (Vec<f64>, i64) = match yconst {
            Some(r) => (vec![r.begin, r.end], 0),
            None => (f.yparams.iter().flat_map(|r| r.iter().cloned()).collect(), if f.yparams.is_empty() { 0 } else { n as i64 }),
        };
        let pts: Vec<f64> = config.points.iter().flat_map(|p| [p.0, p.1]).collect();
        let cfg = ffi::gkq_config {
            algo: self.algo,
            flags: 0,
            tol: tol_to_ffi(&config.tolerance),
            max_evals: config.max_evals as u64,
            workspace_capacity: 0,
            points: if pts.is_empty() { std::ptr::null() } else { pts.as_ptr() },
            npoints: config.points.len() as i64,
        };
        self.last = BatchResult { estimate: vec![0.0; n], delta: vec![0.0; n], nevals: vec![0; n], status: vec![0; n] };
        let csr = !f.pts_offsets.is_empty();
        let rc = unsafe {
            ffi::gkq_integrate2_batch_host(
                self.h, &cfg, fid, yid, n as i64, &xr.begin, &xr.end, 0,
                if fflat.is_empty() { std::ptr::null() } else { fflat.as_ptr() },
                if fflat.is_empty() { 0 } else { n as i64 },
                if yflat.is_empty() { std::ptr::null() } else { yflat.as_ptr() }, ystride,
                if csr { f.pts_offsets.as_ptr() } else { std::ptr::null() },
                if csr { f.pts_xy.as_ptr() } else { std::ptr::null() },
                self.last.estimate.as_mut_ptr(), self.last.delta.as_mut_ptr(),
                self.last.nevals.as_mut_ptr(), self.last.status.as_mut_ptr())
        };
        assert_eq!(rc, 0, "gkq_integrate2_batch_host failed");
        self.last.get(0)
    }
}

macro_rules! algorithm2 {
    ($name:ident, $algo:expr, $doc:expr) => {
        #[doc = $doc]
        pub struct $name(EngineB200);
        impl $name {
            pub fn new(device: i32) -> Result<Self, String> {
                Ok(Self(EngineB200::with_algo(device, $algo)?))
            }
            /// Every nest of the last `integrate` call (the trait returns integral 0, like the 1-D seam).
            pub fn last_batch(&self) -> &BatchResult {
                self.0.last_batch()
            }
            /// Inherent batch method: the whole batch at once, without going through `Integrator2`.
            pub fn integrate_batch(&mut self, f: &DeviceIntegrand2, xr: &Range, yr: Option<&Range>,
                                   config: &IntegrationConfig2) -> &BatchResult {
                self.0.run2(f, xr, yr, config);
                self.0.last_batch()
            }
        }
        /// `Integrator2::with_algorithm(f, QagsB200_2::new(0)?)` over a Rectangle
        /// (double/integrator.rs:80-87 -> this).
        impl Algorithm2<DeviceIntegrand2, Rectangle> for $name {
            fn integrate(&mut self, f: &mut DeviceIntegrand2, range: &Rectangle, config: &IntegrationConfig2) -> IntegrationResult {
                self.0.run2(f, &range.xrange, Some(&range.yrange), config)
            }
        }
    };
}
algorithm2!(Qags2B200, ffi::GKQ_ALGO_QAGS, "replaces gkquad::double::algorithm::QAGS2 (double/algorithm/qags.rs:18-97)");
algorithm2!(Qagp2B200, ffi::GKQ_ALGO_QAGP, "replaces gkquad::double::algorithm::QAGP2 (double/algorithm/qagp.rs:22-139)");
algorithm2!(Auto2B200, ffi::GKQ_ALGO_AUTO, "replaces gkquad::double::algorithm::AUTO2 (double/algorithm/auto.rs:15-36)");

// ---- 1-D inherent batch methods (`integrate_batch`) -----------------------------
impl QagsB200 {
    pub fn integrate_batch(&mut self, f: &DeviceIntegrand, range: &Range, config: &IntegrationConfig) -> &BatchResult {
        self.0.run(f, range, config);
        self.0.last_batch()
    }
}
impl QagpB200 {
    pub fn integrate_batch(&mut self, f: &DeviceIntegrand, range: &Range, config: &IntegrationConfig) -> &BatchResult {
        self.0.run(f, range, config);
        self.0.last_batch()
    }
}
impl AutoB200 {
    pub fn integrate_batch(&mut self, f: &DeviceIntegrand, range: &Range, config: &IntegrationConfig) -> &BatchResult {
        self.0.run(f, range, config);
        self.0.last_batch()
    }
}
