//! `pcp-b200`: libpcp's `Space` with the propagation fixpoint on a B200.
//!
//! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo / rustc: see DESIGN.md 1): this is the source
//! a libpcp maintainer adds next to libpcp to bind `libpcp_b200.so` (include/pcp_b200.h).  It is
//! written against libpcp 0.7.0 @ 9768dd2 and the crates it depends on (`gcollections ^1.4`,
//! `intervallum ^1.2` as `interval`, `trilean ^1.0`); every trait implemented below is cited with
//! the file:line of its definition in the reference.
//!
//! What the types are:
//!   * `GpuVStore<Domain>`  -- a `VStoreConcept` (variable/concept.rs:28-35): owns the engine handle;
//!     the domains live in HBM, a host mirror (`Vec<Domain>`) backs `Index<usize>` / `Iterable` and
//!     is refreshed from the device after every fixpoint.  `Domain` is `Interval<i32>` (VStoreFD,
//!     variable/mod.rs:37) or `IntervalSet<i32>` (VStoreSet, variable/mod.rs:38 -- what FDSpace and
//!     example/src/nqueens.rs use; the engine is then created with PCP_FLAG_INTERVAL_SET).
//!   * `GpuCStore<Domain>`  -- an `IntCStore<GpuVStore<Domain>>` (concept.rs:120-138): keeps the boxed
//!     propagators (for `Collection`, display and `Index`) and lowers each one to a device
//!     descriptor when it is allocated; `Consistency::consistency` (kernel/consistency.rs:17-19)
//!     pushes what is pending and calls `pcp_consistency`.
//!   * `DeviceLowering` (lowering.rs) -- the ONE addition libpcp itself needs: `PropagatorConcept`
//!     (propagation/concept.rs:21-53) and `IntVariable` (concept.rs:79-118) offer no structural
//!     introspection, so both bundles get a `lower()` method.  lowering.rs is written to be
//!     dropped into `src/libpcp/` (it reads private fields of the propagators and views).
//!
//! With these, `search/mod.rs:41-43` becomes
//! ```ignore
//! pub type VStore = pcp_b200::GpuVStore<IntervalSet<i32>>;
//! pub type CStore = pcp_b200::GpuCStore<IntervalSet<i32>>;
//! pub type FDSpace = Space<VStore, CStore, NoRecomputation<VStore, CStore>>;
//! ```
//! and example/src/nqueens.rs, the search combinators (`OneSolution`, `Propagation`, `Brancher`,
//! `BinarySplit`, ...) and `one_solution_engine()` compile and run unchanged: they only use the
//! trait surface implemented here.

pub mod ffi;
pub mod lowering;

use std::cell::RefCell;
use std::fmt::{Debug, Formatter};
use std::ops::Index;
use std::rc::Rc;
use std::slice;

use gcollections::kind::*;
use gcollections::ops::*;
use interval::interval::Interval;
use interval::interval_set::IntervalSet;
use interval::ops::Range;
use pcp::concept::*;
use pcp::kernel::*;
use pcp::model::Model;
use pcp::propagation::concept::PropagatorConcept;
use pcp::propagation::events::FDEvent;
use pcp::term::identity::Identity;
use pcp::variable::ops::{Iterable, MonotonicUpdate};
use trilean::SKleene;

use ffi::*;
pub use lowering::{Desc, DeviceLowering};

// ---------------------------------------------------------------------------------------------
// The domain types the device knows: how a domain is uploaded and how it is rebuilt from what
// `pcp_domains_read` / `pcp_domains_read_bits` return.
// ---------------------------------------------------------------------------------------------
pub trait DeviceDomain: IntDomain<Item = i32> + 'static {
    /// `PCP_FLAG_*` the engine is created with.
    const FLAGS: u32;
    /// Rebuild the domain of one variable from its bounds and (IntervalSet) its bit window.
    fn from_device(lo: i32, hi: i32, base: i32, bits: &[u32]) -> Self;
    /// Whether `from_device` needs the bit window.
    const NEEDS_BITS: bool;
}

impl DeviceDomain for Interval<i32> {
    const FLAGS: u32 = 0;
    const NEEDS_BITS: bool = false;
    fn from_device(lo: i32, hi: i32, _base: i32, _bits: &[u32]) -> Self {
        Interval::new(lo, hi)
    }
}

impl DeviceDomain for IntervalSet<i32> {
    const FLAGS: u32 = PCP_FLAG_INTERVAL_SET;
    const NEEDS_BITS: bool = true;
    fn from_device(lo: i32, hi: i32, base: i32, bits: &[u32]) -> Self {
        // maximal runs of set bits between the bounds, joined left to right
        let has = |v: i32| {
            let o = (v - base) as usize;
            (bits[o >> 5] >> (o & 31)) & 1 == 1
        };
        let mut set = IntervalSet::empty();
        let mut v = lo;
        while v <= hi {
            if !has(v) {
                v += 1;
                continue;
            }
            let mut u = v;
            while u < hi && has(u + 1) {
                u += 1;
            }
            set = set.union(&IntervalSet::new(v, u));
            v = u + 1;
        }
        set
    }
}

// ---------------------------------------------------------------------------------------------
// The engine handle.  One handle per (vstore, cstore) pair, owned by the variable store and
// shared with its frozen form; not `Send` (libpcp's stores are not either: SURVEY 2.1).
// ---------------------------------------------------------------------------------------------
pub struct Engine {
    raw: *mut PcpEngine,
}

impl Engine {
    fn new(flags: u32) -> Engine {
        let cfg = PcpConfig { device: 0, flags, max_labels: 1 << 16, tail_limit: 0 };
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { pcp_engine_create(&cfg, &mut raw) };
        assert!(rc == 0, "pcp_engine_create failed: no sm_100 device (there is no CPU fallback)");
        Engine { raw }
    }
    /// PCP_ERR_INVALID is the C-ABI image of the reference's `assert!` panics (include/pcp_b200.h).
    fn check(&self, rc: i32) {
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(pcp_last_error(self.raw)) };
            panic!("pcp_b200 ({}): {}", rc, msg.to_string_lossy());
        }
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { pcp_engine_destroy(self.raw) }
    }
}

// ---------------------------------------------------------------------------------------------
// GpuVStore: variable/store.rs:29-181 with the domains in HBM.
// ---------------------------------------------------------------------------------------------
pub struct GpuVStore<Domain> {
    engine: Rc<Engine>,
    /// Host mirror: what `Index<usize>` and `Iterable::iter` hand out.  Valid iff `!stale`.
    mirror: RefCell<Vec<Domain>>,
    stale: RefCell<bool>,
}

impl<Domain: DeviceDomain> GpuVStore<Domain> {
    /// Refresh the mirror from the device (after a fixpoint, a restore, an update).
    fn sync(&self) {
        if !*self.stale.borrow() {
            return;
        }
        let n = self.mirror.borrow().len();
        let (mut lo, mut hi) = (vec![0i32; n], vec![0i32; n]);
        self.engine.check(unsafe { pcp_domains_read(self.engine.raw, 0, n as i32, lo.as_mut_ptr(), hi.as_mut_ptr()) });
        let (mut base, mut words, mut bits) = (0i32, 0usize, Vec::new());
        if Domain::NEEDS_BITS && n > 0 {
            base = *lo.iter().min().unwrap();
            words = ((*hi.iter().max().unwrap() - base) as usize) / 32 + 1;
            bits = vec![0u32; n * words];
            self.engine.check(unsafe {
                pcp_domains_read_bits(self.engine.raw, 0, n as i32, base, words as i32, bits.as_mut_ptr())
            });
        }
        let mut m = self.mirror.borrow_mut();
        for i in 0..n {
            let w = if Domain::NEEDS_BITS { &bits[i * words..(i + 1) * words] } else { &bits[..] };
            m[i] = Domain::from_device(lo[i], hi[i], base, w);
        }
        *self.stale.borrow_mut() = false;
    }
    fn invalidate(&self) {
        *self.stale.borrow_mut() = true;
    }
    pub(crate) fn engine(&self) -> &Rc<Engine> {
        &self.engine
    }
}

impl<Domain> Collection for GpuVStore<Domain> {
    type Item = Domain; // variable/store.rs:35-40
}

impl<Domain> AssociativeCollection for GpuVStore<Domain> {
    type Location = Identity<Domain>; // variable/store.rs:42-47
}

impl<Domain: DeviceDomain> Empty for GpuVStore<Domain> {
    fn empty() -> Self {
        // variable/store.rs:77-84
        GpuVStore { engine: Rc::new(Engine::new(Domain::FLAGS)), mirror: RefCell::new(vec![]), stale: RefCell::new(false) }
    }
}

impl<Domain> Cardinality for GpuVStore<Domain> {
    type Size = usize;
    fn size(&self) -> usize {
        self.mirror.borrow().len() // variable/store.rs:109-118
    }
}

impl<Domain: DeviceDomain> Iterable for GpuVStore<Domain> {
    fn iter(&self) -> slice::Iter<'_, Domain> {
        // variable/store.rs:120-127.  The mirror is only replaced element-wise in `sync`, never
        // reallocated while a borrow is out (libpcp's stores hand out `&Domain` the same way).
        self.sync();
        unsafe { (*self.mirror.as_ptr()).iter() }
    }
}

impl<Domain: DeviceDomain> Alloc for GpuVStore<Domain> {
    fn alloc(&mut self, dom: Domain) -> Identity<Domain> {
        // variable/store.rs:129-141: `assert!(!dom.is_empty())` -> PCP_ERR_INVALID -> panic
        assert!(!dom.is_empty());
        let (lo, hi) = (dom.lower(), dom.upper());
        let mut first = 0i32;
        self.engine.check(unsafe { pcp_vars_alloc(self.engine.raw, &lo, &hi, 1, &mut first) });
        // an IntervalSet allocated with holes: remove them value by value is not needed by any
        // libpcp caller (every model allocates `IntervalSet::new(l, u)`); refuse it loudly
        assert!(dom.size() as i64 == hi as i64 - lo as i64 + 1, "alloc of a domain with holes is not supported");
        self.mirror.borrow_mut().push(dom);
        Identity::new(first as usize)
    }
}

impl<Domain: DeviceDomain> MonotonicUpdate for GpuVStore<Domain> {
    fn update(&mut self, loc: &Identity<Domain>, dom: Domain) -> bool {
        // variable/store.rs:143-167.  The C ABI takes the hull; on IntervalSet engines the result
        // is `cur /\ [lo, hi]` (include/pcp_b200.h) -- exact for every update libpcp's search layer
        // performs (it only narrows through propagators, which run on the device).
        if dom.is_empty() {
            return false;
        }
        let mut ok = 0i32;
        self.engine.check(unsafe { pcp_var_update(self.engine.raw, loc.index() as i32, dom.lower(), dom.upper(), &mut ok) });
        self.invalidate();
        ok != 0
    }
}

impl<Domain: DeviceDomain> Index<usize> for GpuVStore<Domain> {
    type Output = Domain;
    fn index(&self, index: usize) -> &Domain {
        // variable/store.rs:169-182
        self.sync();
        assert!(index < self.size(), "Variable not registered in the store. Variable index must be obtained with `alloc`.");
        unsafe { &(*self.mirror.as_ptr())[index] }
    }
}

impl<Domain: Debug> Debug for GpuVStore<Domain> {
    fn fmt(&self, f: &mut Formatter) -> std::fmt::Result {
        f.debug_struct("GpuVStore").field("domains", &*self.mirror.borrow()).finish()
    }
}

impl<Domain: DeviceDomain + std::fmt::Display> DisplayStateful<Model> for GpuVStore<Domain> {
    fn display(&self, model: &Model) {
        // variable/store.rs:205-223 (one line per variable instead of the matrix layout)
        self.sync();
        for (i, d) in self.mirror.borrow().iter().enumerate() {
            println!("{} = {}", model.var_name(i), d);
        }
    }
}

impl<Domain: DeviceDomain> ImmutableMemoryConcept for GpuVStore<Domain> {} // variable/memory/concept.rs:23-33
impl<Domain: DeviceDomain + std::fmt::Display> VStoreConcept for GpuVStore<Domain> {} // variable/concept.rs:28-35

/// kernel/restoration.rs:15-30 for the variable store.  A label is the engine's label (domains,
/// number of propagators and `active` set as of `label()`): plain `u64`, clonable, valid until a
/// restore to an older one -- the contract of `TrailMemory`'s marks (trail_memory.rs:148-160).
pub struct FrozenGpuVStore<Domain> {
    store: GpuVStore<Domain>,
}

impl<Domain: DeviceDomain> Freeze for GpuVStore<Domain> {
    type FrozenState = FrozenGpuVStore<Domain>;
    fn freeze(self) -> FrozenGpuVStore<Domain> {
        FrozenGpuVStore { store: self } // variable/store.rs:239-247
    }
}

impl<Domain: DeviceDomain> Snapshot for FrozenGpuVStore<Domain> {
    type Label = u64;
    type State = GpuVStore<Domain>;
    fn label(&mut self) -> u64 {
        // variable/store.rs:276-278.  Right after an Unknown fixpoint the engine already holds the
        // label copy (the kernel epilogue leaves it in the next slot): no device work here.
        let mut l = 0u64;
        self.store.engine.check(unsafe { pcp_label(self.store.engine.raw, &mut l) });
        l
    }
    fn restore(self, label: u64) -> GpuVStore<Domain> {
        // variable/store.rs:280-283.  The copy rides inside the next fixpoint launch.
        self.store.engine.check(unsafe { pcp_restore(self.store.engine.raw, label) });
        self.store.invalidate();
        self.store
    }
}

// ---------------------------------------------------------------------------------------------
// GpuCStore: propagation/store.rs:32-324 with the propagators, the reactor, the scheduler and the
// fixpoint loop on the device.
// ---------------------------------------------------------------------------------------------
pub struct GpuCStore<Domain> {
    /// The boxed propagators, kept for `Collection` / `Index` / display (store.rs:33).
    propagators: Vec<Formula<GpuVStore<Domain>>>,
    /// Device descriptors of `propagators[pushed..]`, not yet handed to the engine: `Alloc::alloc`
    /// has no access to the variable store (hence to the engine), `consistency` has.
    pending: Vec<Desc>,
    pushed: usize,
}

impl<Domain: DeviceDomain> GpuCStore<Domain> {
    fn push_pending(&mut self, engine: &Engine) {
        for d in self.pending.drain(..) {
            let mut idx = 0i32;
            let rc = match d {
                Desc::Prop { kind, views } => {
                    // Sum::new (term/sum.rs:28-32): multi-term sums are registered first
                    let ops: Vec<PcpOperand> = views
                        .iter()
                        .map(|v| match v {
                            lowering::ViewDesc::Operand(o) => *o,
                            lowering::ViewDesc::Sum { terms, off } if terms.len() == 1 => {
                                PcpOperand { var: terms[0].var, off: terms[0].off + off }
                            }
                            lowering::ViewDesc::Sum { terms, off } => {
                                let mut id = 0i32;
                                engine.check(unsafe { pcp_sum_alloc(engine.raw, terms.as_ptr(), terms.len() as i32, &mut id) });
                                PcpOperand { var: var_sum(id), off: *off }
                            }
                        })
                        .collect();
                    unsafe { pcp_prop_alloc(engine.raw, kind, ops.as_ptr(), ops.len() as i32, &mut idx) }
                }
                Desc::Formula { words } => unsafe { pcp_formula_alloc(engine.raw, words.as_ptr(), words.len() as i32, &mut idx) },
            };
            engine.check(rc); // PCP_ERR_UNSUPPORTED: a propagator without a device lowering (no CPU fallback)
            assert_eq!(idx as usize, self.pushed, "device and host propagator indices diverged");
            self.pushed += 1;
        }
    }
}

impl<Domain> Collection for GpuCStore<Domain> {
    type Item = Formula<GpuVStore<Domain>>; // store.rs:55-57
}

impl<Domain> AssociativeCollection for GpuCStore<Domain> {
    type Location = usize; // store.rs:59-61
}

impl<Domain> Cardinality for GpuCStore<Domain> {
    type Size = usize;
    fn size(&self) -> usize {
        self.propagators.len() // store.rs:63-69
    }
}

impl<Domain> Empty for GpuCStore<Domain> {
    fn empty() -> Self {
        GpuCStore { propagators: vec![], pending: vec![], pushed: 0 } // store.rs:39-53
    }
}

impl<Domain: DeviceDomain> Alloc for GpuCStore<Domain> {
    fn alloc(&mut self, p: Self::Item) -> usize {
        // store.rs:223-230.  Lowered here so that an un-lowerable propagator fails where it is posted.
        let d = p.lower().expect("propagator has no device lowering (pcp_b200 has no CPU fallback)");
        let idx = self.propagators.len();
        self.propagators.push(p);
        self.pending.push(d);
        idx
    }
}

impl<Domain> Index<usize> for GpuCStore<Domain> {
    type Output = Formula<GpuVStore<Domain>>;
    fn index(&self, index: usize) -> &Self::Output {
        &self.propagators[index] // store.rs:210-215
    }
}

impl<Domain: DeviceDomain> Consistency<GpuVStore<Domain>> for GpuCStore<Domain> {
    fn consistency(&mut self, vstore: &mut GpuVStore<Domain>) -> SKleene {
        // store.rs:247-257 = one `pcp_consistency` = one launch of the device fixpoint
        let engine = vstore.engine().clone();
        self.push_pending(&engine);
        let mut status = 0i32;
        engine.check(unsafe { pcp_consistency(engine.raw, &mut status, std::ptr::null_mut()) });
        vstore.invalidate();
        match status {
            PCP_FALSE => SKleene::False,
            PCP_TRUE => SKleene::True,
            _ => SKleene::Unknown,
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Several nodes at once (no counterpart in libpcp, which is single-threaded; sibling subtrees are
// independent: search/branching/branch.rs:36-55).  `fork_space` = `Space::clone` for a sibling
// subtree without a second copy of the model on the device (pcp_engine_fork); `consistency_many` =
// `Consistency::consistency` of several such pairs, served by ONE launch when the pairs are forks of
// one model (pcp_consistency_batch: a group of CTAs per pair).
// ---------------------------------------------------------------------------------------------
pub fn fork_space<Domain: DeviceDomain>(vstore: &GpuVStore<Domain>, cstore: &mut GpuCStore<Domain>) -> (GpuVStore<Domain>, GpuCStore<Domain>) {
    let engine = vstore.engine().clone();
    cstore.push_pending(&engine);
    let mut raw = std::ptr::null_mut();
    engine.check(unsafe { pcp_engine_fork(engine.raw, &mut raw) });
    vstore.sync();
    let child = GpuVStore { engine: Rc::new(Engine { raw }), mirror: RefCell::new(vstore.mirror.borrow().clone()), stale: RefCell::new(false) };
    (child, GpuCStore { propagators: cstore.propagators.iter().map(|p| p.bclone()).collect(), pending: vec![], pushed: cstore.pushed })
}

pub fn consistency_many<Domain: DeviceDomain>(spaces: &mut [(&mut GpuVStore<Domain>, &mut GpuCStore<Domain>)]) -> Vec<SKleene> {
    let mut raws = Vec::with_capacity(spaces.len());
    for (v, c) in spaces.iter_mut() {
        let engine = v.engine().clone();
        c.push_pending(&engine);
        raws.push(engine.raw);
    }
    let mut status = vec![0i32; spaces.len()];
    let rc = unsafe { pcp_consistency_batch(raws.as_ptr(), raws.len() as i32, status.as_mut_ptr(), std::ptr::null_mut()) };
    for (v, _) in spaces.iter_mut() {
        if rc != 0 {
            v.engine().check(rc);
        }
        v.invalidate();
    }
    status.iter().map(|&s| match s { PCP_FALSE => SKleene::False, PCP_TRUE => SKleene::True, _ => SKleene::Unknown }).collect()
}

impl<Domain> Clone for GpuCStore<Domain> {
    fn clone(&self) -> Self {
        // store.rs:260-272 clones the boxed propagators; the device state belongs to the engine of
        // the variable store this clone is used with.  Only `Debugger` / display code clones.
        GpuCStore { propagators: self.propagators.iter().map(|p| p.bclone()).collect(), pending: vec![], pushed: self.pushed }
    }
}

impl<Domain> DisplayStateful<Model> for GpuCStore<Domain> {
    fn display(&self, model: &Model) {
        // store.rs:105-116
        for p in &self.propagators {
            p.display(model);
            println!();
        }
    }
}

impl<Domain> DisplayStateful<(Model, GpuVStore<Domain>)> for GpuCStore<Domain> {
    fn display(&self, (model, _vstore): &(Model, GpuVStore<Domain>)) {
        DisplayStateful::<Model>::display(self, model) // store.rs:87-103 (without the per-status grouping)
    }
}

/// store.rs:286-324: the label of the constraint store is the number of propagators; the `active`
/// bit set it pairs with in the reference lives on the device and is restored by the variable
/// store's label (`NoRecomputation` always restores the two together,
/// search/recomputation/no_recomputation.rs:49-61).
pub struct FrozenGpuCStore<Domain> {
    cstore: GpuCStore<Domain>,
}

impl<Domain: DeviceDomain> Freeze for GpuCStore<Domain> {
    type FrozenState = FrozenGpuCStore<Domain>;
    fn freeze(self) -> FrozenGpuCStore<Domain> {
        assert!(self.pending.is_empty(), "freeze() before the posted propagators reached a fixpoint");
        FrozenGpuCStore { cstore: self }
    }
}

impl<Domain: DeviceDomain> Snapshot for FrozenGpuCStore<Domain> {
    type Label = usize;
    type State = GpuCStore<Domain>;
    fn label(&mut self) -> usize {
        self.cstore.propagators.len() // store.rs:315-317
    }
    fn restore(mut self, label: usize) -> GpuCStore<Domain> {
        // store.rs:319-323: truncate; the engine truncates its own list in pcp_restore
        self.cstore.propagators.truncate(label);
        self.cstore.pushed = label;
        self.cstore
    }
}

// `IntCStore<GpuVStore<Domain>>` (concept.rs:120-138) is a blanket impl over exactly the traits
// implemented above: Alloc + Empty + Clone + Freeze + DisplayStateful<Model> +
// DisplayStateful<(Model, VStore)> + Collection<Item = Formula<VStore>> + Consistency<VStore>.

#[allow(dead_code)]
fn _assert_bundles() {
    fn is_cstore<V: Collection, C: IntCStore<V>>() {}
    fn is_vstore<V: VStoreConcept>() {}
    is_vstore::<GpuVStore<Interval<i32>>>();
    is_vstore::<GpuVStore<IntervalSet<i32>>>();
    is_cstore::<GpuVStore<Interval<i32>>, GpuCStore<Interval<i32>>>();
    is_cstore::<GpuVStore<IntervalSet<i32>>, GpuCStore<IntervalSet<i32>>>();
    let _ = FDEvent::Bound;
    let _: Option<Box<dyn PropagatorConcept<GpuVStore<Interval<i32>>, FDEvent>>> = None;
    let _ = <Interval<i32> as Range>::new(0, 1);
}
