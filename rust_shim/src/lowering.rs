//! `DeviceLowering`: the one addition libpcp itself needs (NOT COMPILED HERE: no Rust toolchain).
//!
//! `PropagatorConcept` (propagation/concept.rs:21-53) and `IntVariable` (concept.rs:79-118) are
//! bundles of behaviour -- propagate / is_subsumed / dependencies / display / not / bclone, and
//! read / update / dependencies / display / bclone -- with no way to look inside a boxed
//! propagator or view.  The device needs the structure (kind + operands), so both bundles get
//! one more supertrait with one method.  This file is written to live INSIDE libpcp (as
//! `src/libpcp/device_lowering.rs`, `pub mod device_lowering;` in lib.rs): the impls read the
//! private fields of the propagators and views.  The two edits to existing libpcp lines are
//!
//! ```ignore
//! // concept.rs:79-86        pub trait IntVariable_<VStore>: ViewDependencies<FDEvent> + ... + ViewLowering
//! // propagation/concept.rs:21-29   pub trait PropagatorConcept_<VStore, Event>: Propagator<VStore> + ... + DeviceLowering
//! ```
//!
//! Everything that returns `None` is rejected when it is allocated into a `GpuCStore`
//! ("propagator has no device lowering"): there is no CPU fallback on the fixpoint path.

use crate::ffi::*;

/// A view as the device sees it (include/pcp_b200.h, `pcp_operand`): `var >= 0` Identity (+ offset,
/// nested Additions fold), `var == -1` Constant(off), `var <= -2` a Sum view, whose terms travel
/// beside it because the engine numbers Sum views itself (`pcp_sum_alloc`).
#[derive(Clone, Debug)]
pub enum ViewDesc {
    Operand(PcpOperand),
    Sum { terms: Vec<PcpOperand>, off: i32 },
}

pub trait ViewLowering {
    fn lower_view(&self) -> Option<ViewDesc> {
        None
    }
}

/// What one allocated propagator becomes: a descriptor of one of the fixed kinds
/// (`pcp_prop_alloc`) or a formula tree (`pcp_formula_alloc`).
#[derive(Clone, Debug)]
pub enum Desc {
    /// `views` still carry multi-term Sum views as their terms: `GpuCStore::push_pending` registers
    /// them with `pcp_sum_alloc` and passes `PCP_VAR_SUM(id)` operands.
    Prop { kind: i32, views: Vec<ViewDesc> },
    Formula { words: Vec<i32> },
}

pub trait DeviceLowering {
    fn lower(&self) -> Option<Desc> {
        None
    }
    /// The same propagator as a node of a formula tree (prefix words); `None` = cannot be a child
    /// of Conjunction / Disjunction on the device.
    fn lower_node(&self) -> Option<Vec<i32>> {
        None
    }
}

fn operand(v: &ViewDesc) -> Option<PcpOperand> {
    match v {
        ViewDesc::Operand(o) => Some(*o),
        ViewDesc::Sum { terms, off } if terms.len() == 1 => Some(PcpOperand { var: terms[0].var, off: terms[0].off + off }), // term/sum.rs:62-64
        _ => None, // multi-term sums need pcp_sum_alloc first: done by GpuCStore::push_pending through `ViewDesc::Sum`
    }
}

/// The operands of a fixed-kind propagator, and the same as a formula-tree leaf (`None` when an
/// operand is a multi-term Sum: trees hold plain operands only).
fn leaf(kind: i32, views: &[Option<ViewDesc>]) -> Option<(Vec<ViewDesc>, Option<Vec<i32>>)> {
    let mut out = vec![];
    let mut words = Some(vec![PCP_F_LEAF + kind]);
    for v in views {
        let v = v.clone()?;
        match (operand(&v), words.as_mut()) {
            (Some(o), Some(w)) => {
                w.push(o.var);
                w.push(o.off);
            }
            _ => words = None,
        }
        out.push(v);
    }
    Some((out, words))
}

// ---- views (term/) -----------------------------------------------------------------------------
mod views {
    use super::*;
    use concept::*;
    use gcollections::kind::*;
    use term::addition::Addition;
    use term::constant::Constant;
    use term::identity::Identity;
    use term::sum::Sum;

    impl<Domain> ViewLowering for Identity<Domain> {
        fn lower_view(&self) -> Option<ViewDesc> {
            Some(ViewDesc::Operand(PcpOperand { var: self.index() as i32, off: 0 })) // term/identity.rs:23-39
        }
    }

    impl<VStore> ViewLowering for Addition<VStore>
    where
        VStore: VStoreConcept,
        VStore::Item: Collection<Item = i32>,
    {
        fn lower_view(&self) -> Option<ViewDesc> {
            // term/addition.rs:24-31: x + v; nested additions fold into one offset
            match self.x.lower_view()? {
                ViewDesc::Operand(o) => Some(ViewDesc::Operand(PcpOperand { var: o.var, off: o.off + self.v })),
                ViewDesc::Sum { terms, off } => Some(ViewDesc::Sum { terms, off: off + self.v }),
            }
        }
    }

    impl ViewLowering for Constant<i32> {
        fn lower_view(&self) -> Option<ViewDesc> {
            Some(ViewDesc::Operand(PcpOperand { var: PCP_VAR_CONSTANT, off: self.value })) // term/constant.rs:24-32
        }
    }

    impl<VStore: Collection> ViewLowering for Sum<VStore> {
        fn lower_view(&self) -> Option<ViewDesc> {
            // term/sum.rs:23-31: the terms must themselves be plain operands
            let mut terms = vec![];
            for v in &self.vars {
                match v.lower_view()? {
                    ViewDesc::Operand(o) => terms.push(o),
                    ViewDesc::Sum { .. } => return None, // nested sums: no device lowering
                }
            }
            Some(ViewDesc::Sum { terms, off: 0 })
        }
    }
}

// ---- propagators (propagators/, logic/) ---------------------------------------------------------
mod props {
    use super::*;
    use gcollections::kind::*;
    use logic::boolean::Boolean;
    use logic::boolean_neg::BooleanNeg;
    use logic::conjunction::Conjunction;
    use logic::disjunction::Disjunction;
    use propagators::all_equal::AllEqual;
    use propagators::cmp::x_eq_y::XEqY;
    use propagators::cmp::x_eq_y_mul_z::XEqYMulZ;
    use propagators::cmp::x_eq_y_plus_z::XEqYPlusZ;
    use propagators::cmp::x_greater_y_plus_z::XGreaterYPlusZ;
    use propagators::cmp::x_less_y::XLessY;
    use propagators::cmp::x_less_y_plus_z::XLessYPlusZ;
    use propagators::cmp::x_neq_y::XNeqY;
    use propagators::distinct::Distinct;

    macro_rules! fixed_kind {
        ($ty:ident, $kind:expr, $($field:ident),+) => {
            impl<VStore: Collection> DeviceLowering for $ty<VStore> {
                fn lower(&self) -> Option<Desc> {
                    let (views, _) = leaf($kind, &[$(self.$field.lower_view()),+])?;
                    Some(Desc::Prop { kind: $kind, views })
                }
                fn lower_node(&self) -> Option<Vec<i32>> {
                    leaf($kind, &[$(self.$field.lower_view()),+])?.1
                }
            }
        };
    }
    fixed_kind!(XLessY, PCP_X_LESS_Y, x, y); // cmp/x_less_y.rs:28-31 (also x_greater_y / x_geq_y / x_leq_y, cmp/mod.rs:34-60)
    fixed_kind!(XNeqY, PCP_X_NEQ_Y, x, y); // cmp/x_neq_y.rs:27-30
    fixed_kind!(XEqY, PCP_X_EQ_Y, x, y); // cmp/x_eq_y.rs:28-31
    fixed_kind!(XGreaterYPlusZ, PCP_X_GREATER_Y_PLUS_Z, x, y, z); // cmp/x_greater_y_plus_z.rs:29-33
    fixed_kind!(XLessYPlusZ, PCP_X_LESS_Y_PLUS_Z, x, y, z); // cmp/x_less_y_plus_z.rs:29-33

    impl<VStore: Collection> DeviceLowering for XEqYMulZ<VStore> {
        fn lower(&self) -> Option<Desc> {
            // cmp/x_eq_y_mul_z.rs:31-35; not a formula-tree leaf (its not() is unimplemented!())
            let (views, _) = leaf(PCP_X_EQ_Y_MUL_Z, &[self.x.lower_view(), self.y.lower_view(), self.z.lower_view()])?;
            Some(Desc::Prop { kind: PCP_X_EQ_Y_MUL_Z, views })
        }
    }

    impl<VStore: Collection> DeviceLowering for XEqYPlusZ<VStore> {
        // cmp/x_eq_y_plus_z.rs:26-49: geq = XGreaterYPlusZ(x + 1, y, z), leq = XLessYPlusZ(x - 1, y, z);
        // the device kind takes (x, y, z) and fuses the two halves
        fn lower(&self) -> Option<Desc> {
            let x = match self.geq.x.lower_view()? {  // undo the Addition(x, 1) of x_geq_y_plus_z (cmp/mod.rs:70-76)
                ViewDesc::Operand(o) => ViewDesc::Operand(PcpOperand { var: o.var, off: o.off - 1 }),
                ViewDesc::Sum { terms, off } => ViewDesc::Sum { terms, off: off - 1 },
            };
            let (views, _) = leaf(PCP_X_EQ_Y_PLUS_Z, &[Some(x), self.geq.y.lower_view(), self.geq.z.lower_view()])?;
            Some(Desc::Prop { kind: PCP_X_EQ_Y_PLUS_Z, views })
        }
        fn lower_node(&self) -> Option<Vec<i32>> {
            match self.lower()? {
                Desc::Prop { views, .. } => {
                    let mut w = vec![PCP_F_LEAF + PCP_X_EQ_Y_PLUS_Z];
                    for v in &views {
                        let o = operand(v)?;
                        w.push(o.var);
                        w.push(o.off);
                    }
                    Some(w)
                }
                _ => None,
            }
        }
    }

    fn nary<VStore: Collection>(kind: i32, vars: &[Var<VStore>]) -> Option<Desc> {
        let mut views = vec![];
        for v in vars {
            views.push(ViewDesc::Operand(operand(&v.lower_view()?)?));  // (no Sum operands inside n-ary propagators)
        }
        Some(Desc::Prop { kind, views })
    }
    impl<VStore: Collection> DeviceLowering for Distinct<VStore> {
        fn lower(&self) -> Option<Desc> {
            nary(PCP_DISTINCT, &self.vars) // propagators/distinct.rs:48-51: one n-ary descriptor instead of n(n-1)/2 pairs
        }
    }
    impl<VStore: Collection> DeviceLowering for AllEqual<VStore> {
        fn lower(&self) -> Option<Desc> {
            nary(PCP_ALL_EQUAL, &self.vars) // propagators/all_equal.rs:26-29
        }
    }

    fn connective<VStore>(tag: i32, fs: &[Formula<VStore>]) -> Option<Vec<i32>> {
        let mut w = vec![tag, fs.len() as i32];
        for f in fs {
            w.extend(f.lower_node()?);
        }
        Some(w)
    }
    impl<VStore> DeviceLowering for Conjunction<VStore> {
        // logic/conjunction.rs:25-27.  (The reference's not() has already been applied when a
        // negated formula reaches alloc -- `g.not()` builds the complemented tree, logic/mod.rs:30-35 --
        // so PCP_F_NOT is never needed from this side.)
        fn lower(&self) -> Option<Desc> {
            Some(Desc::Formula { words: self.lower_node()? })
        }
        fn lower_node(&self) -> Option<Vec<i32>> {
            connective(PCP_F_CONJUNCTION, &self.fs)
        }
    }
    impl<VStore> DeviceLowering for Disjunction<VStore> {
        fn lower(&self) -> Option<Desc> {
            Some(Desc::Formula { words: self.lower_node()? }) // logic/disjunction.rs:25-27
        }
        fn lower_node(&self) -> Option<Vec<i32>> {
            connective(PCP_F_DISJUNCTION, &self.fs)
        }
    }
    impl<VStore> DeviceLowering for Boolean<VStore> {
        fn lower(&self) -> Option<Desc> {
            Some(Desc::Formula { words: self.lower_node()? }) // logic/boolean.rs:29-31
        }
        fn lower_node(&self) -> Option<Vec<i32>> {
            let o = operand(&self.var.lower_view()?)?;
            Some(vec![PCP_F_BOOLEAN, o.var, o.off])
        }
    }
    impl<VStore> DeviceLowering for BooleanNeg<VStore> {
        fn lower(&self) -> Option<Desc> {
            Some(Desc::Formula { words: self.lower_node()? }) // logic/boolean_neg.rs:30-32
        }
        fn lower_node(&self) -> Option<Vec<i32>> {
            let o = operand(&self.b.var.lower_view()?)?;
            Some(vec![PCP_F_BOOLEAN_NEG, o.var, o.off])
        }
    }
}
