//! `GpuCStore`: a constraint store whose `Consistency::consistency` runs on a B200.
//!
//! Source only -- this image has no Rust toolchain; it documents, in the reference's own
//! language, exactly what the C ABI of `include/pcp_b200.h` replaces:
//!   * `Store::alloc`        (libpcp `propagation/store.rs:223-230`)  -> `pcp_prop_alloc`
//!   * `Store::consistency`  (`propagation/store.rs:247-257`)         -> `pcp_consistency`
//!   * `FrozenStore::{label, restore}` (`propagation/store.rs:312-323`) together with the
//!     variable store's snapshot (`variable/store.rs:274-283`)       -> `pcp_label/pcp_restore`
//!   * `VStore::alloc / update / index` (`variable/store.rs:135-181`) -> `pcp_vars_alloc`,
//!     `pcp_var_update`, `pcp_domains_read`
//! Model code (`example/src/nqueens.rs`) is unchanged: it allocates into `space.vstore` /
//! `space.cstore` as before; only the `CStore`/`VStore` type aliases of `search/mod.rs:41-43`
//! point at the types below.

use std::os::raw::{c_char, c_int};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PcpOperand { pub var: i32, pub off: i32 }

#[repr(C)]
pub struct PcpConfig { pub device: i32, pub flags: u32, pub max_labels: u32, pub tail_limit: u32 }

#[repr(C)]
#[derive(Default)]
pub struct PcpStats { pub propagations: u64, pub iterations: u32, pub active_props: u32, pub kernel_ms: f32, pub reserved: u32 }

pub enum PcpEngine {}

pub const PCP_X_LESS_Y: i32 = 0;
pub const PCP_X_NEQ_Y: i32 = 1;
pub const PCP_X_EQ_Y: i32 = 2;
pub const PCP_X_GREATER_Y_PLUS_Z: i32 = 3;
pub const PCP_X_LESS_Y_PLUS_Z: i32 = 4;
pub const PCP_X_EQ_Y_PLUS_Z: i32 = 5;
pub const PCP_DISTINCT: i32 = 6;
pub const PCP_DISJ2_X_EQ_Y_PLUS_Z: i32 = 7;
pub const PCP_X_EQ_Y_MUL_Z: i32 = 8;
pub const PCP_ALL_EQUAL: i32 = 9;

pub const PCP_FLAG_INCREMENTAL: u32 = 1;
pub const PCP_FLAG_HOST_SEARCH: u32 = 2;

extern "C" {
    pub fn pcp_engine_create(cfg: *const PcpConfig, out: *mut *mut PcpEngine) -> c_int;
    pub fn pcp_engine_destroy(e: *mut PcpEngine);
    pub fn pcp_last_error(e: *const PcpEngine) -> *const c_char;
    pub fn pcp_vars_alloc(e: *mut PcpEngine, lo: *const i32, hi: *const i32, n: i32, first: *mut i32) -> c_int;
    pub fn pcp_sum_alloc(e: *mut PcpEngine, terms: *const PcpOperand, n: i32, sum_id: *mut i32) -> c_int;
    pub fn pcp_prop_alloc(e: *mut PcpEngine, kind: i32, ops: *const PcpOperand, n_ops: i32, idx: *mut i32) -> c_int;
    pub fn pcp_consistency(e: *mut PcpEngine, status: *mut i32, stats: *mut PcpStats) -> c_int;
    pub fn pcp_domains_read(e: *mut PcpEngine, first: i32, n: i32, lo: *mut i32, hi: *mut i32) -> c_int;
    pub fn pcp_var_update(e: *mut PcpEngine, idx: i32, lo: i32, hi: i32, ok: *mut i32) -> c_int;
    pub fn pcp_label(e: *mut PcpEngine, label: *mut u64) -> c_int;
    pub fn pcp_restore(e: *mut PcpEngine, label: u64) -> c_int;
    pub fn pcp_stream(e: *mut PcpEngine, stream: *mut *mut std::os::raw::c_void) -> c_int;
}

/// What a propagator lowers to.  `PropagatorConcept` (libpcp `propagation/concept.rs:21-53`)
/// offers no structural introspection, so the bundle gets ONE added, defaulted method:
///
/// ```ignore
/// pub trait DeviceLowering { fn lower(&self) -> Option<Desc> { None } }
/// ```
/// implemented for the hot-path propagators (cmp/x_less_y.rs, x_neq_y.rs, x_eq_y.rs,
/// x_greater_y_plus_z.rs, x_less_y_plus_z.rs, x_eq_y_plus_z.rs, x_eq_y_mul_z.rs, distinct.rs, and a
/// `Disjunction` of two `XEqYPlusZ`) and for the four views (`Identity` -> (idx, 0),
/// `Addition` -> inner + v, `Constant` -> (-1, value), `Sum` -> a sum id).  Anything that
/// returns `None` is an error at `alloc` time: there is no CPU fallback on the fixpoint path.
pub struct Desc { pub kind: i32, pub ops: Vec<PcpOperand> }

pub trait DeviceLowering {
    fn lower(&self) -> Option<Desc> { None }
}

/// The engine handle shared by the two stores of a `Space` (`search/space.rs:21-25`).
pub struct Engine { raw: *mut PcpEngine }

impl Engine {
    pub fn new(device: i32) -> Engine {
        let cfg = PcpConfig { device, flags: 0, max_labels: 0, tail_limit: 0 };
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { pcp_engine_create(&cfg, &mut raw) };
        assert!(rc == 0, "pcp_engine_create failed: no sm_100 device (there is no CPU fallback)");
        Engine { raw }
    }
    fn check(&self, rc: c_int) {
        // PCP_ERR_INVALID is the C-ABI image of the reference's assert! panics
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(pcp_last_error(self.raw)) };
            panic!("pcp_b200: {}", msg.to_string_lossy());
        }
    }
}
impl Drop for Engine { fn drop(&mut self) { unsafe { pcp_engine_destroy(self.raw) } } }

/// Constraint store: `Alloc` + `Consistency<VStore>` + `Freeze`/`Snapshot` as required by
/// `IntCStore` (libpcp `concept.rs:120-138`).
pub struct GpuCStore { engine: std::rc::Rc<Engine>, len: usize }

impl GpuCStore {
    /// `Alloc::alloc` (`propagation/store.rs:223-230`).
    pub fn alloc<P: DeviceLowering>(&mut self, p: &P) -> usize {
        let d = p.lower().expect("propagator has no device lowering");
        let mut idx = 0i32;
        let rc = unsafe { pcp_prop_alloc(self.engine.raw, d.kind, d.ops.as_ptr(), d.ops.len() as i32, &mut idx) };
        self.engine.check(rc);
        self.len += 1;
        idx as usize
    }
    /// `Consistency::consistency` (`kernel/consistency.rs:17-19`): -1 False, 0 Unknown, 1 True
    /// maps onto `trilean::SKleene`.
    pub fn consistency(&mut self) -> i32 {
        let mut status = 0i32;
        let mut stats = PcpStats::default();
        let rc = unsafe { pcp_consistency(self.engine.raw, &mut status, &mut stats) };
        self.engine.check(rc);
        status
    }
    /// `Snapshot::label` / `Snapshot::restore` of the (vstore, cstore) pair, as
    /// `NoRecomputation` does (`search/recomputation/no_recomputation.rs:49-61`).
    pub fn label(&mut self) -> u64 {
        let mut l = 0u64;
        let rc = unsafe { pcp_label(self.engine.raw, &mut l) };
        self.engine.check(rc);
        l
    }
    pub fn restore(&mut self, label: u64) {
        let rc = unsafe { pcp_restore(self.engine.raw, label) };
        self.engine.check(rc);
    }
}

/// Variable store mirror: domains live on the device, reads go through `pcp_domains_read`
/// (used by `FirstSmallestVar` / `Brancher`, `search/branching/first_smallest_var.rs:30-39`).
pub struct GpuVStore { engine: std::rc::Rc<Engine>, len: usize }

impl GpuVStore {
    /// `VStore::alloc` (`variable/store.rs:135-140`).
    pub fn alloc(&mut self, lo: i32, hi: i32) -> usize {
        let mut first = 0i32;
        let rc = unsafe { pcp_vars_alloc(self.engine.raw, &lo, &hi, 1, &mut first) };
        self.engine.check(rc);
        self.len += 1;
        first as usize
    }
    /// `Index<usize>` (`variable/store.rs:175-181`).
    pub fn read(&self, idx: usize) -> (i32, i32) {
        let (mut lo, mut hi) = (0i32, 0i32);
        let rc = unsafe { pcp_domains_read(self.engine.raw, idx as i32, 1, &mut lo, &mut hi) };
        self.engine.check(rc);
        (lo, hi)
    }
    /// `MonotonicUpdate::update` (`variable/store.rs:151-166`).
    pub fn update(&mut self, idx: usize, lo: i32, hi: i32) -> bool {
        let mut ok = 0i32;
        let rc = unsafe { pcp_var_update(self.engine.raw, idx as i32, lo, hi, &mut ok) };
        self.engine.check(rc);
        ok != 0
    }
}
