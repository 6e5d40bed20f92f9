// pcp_device.cuh -- sm_100a device side of the propagation engine.
//
// One persistent kernel (`pcp_fixpoint_kernel`) runs the whole propagation fixpoint of a
// search node -- the device equivalent of Store::consistency
// (reference src/libpcp/propagation/store.rs:247-257): one CTA per SM, a device-wide
// barrier between iterations, quiescence detected by the last CTA to arrive.
//
//   prologue      Snapshot::restore + Store::alloc of the few propagators posted since the
//                 last call (CTA 0), overlapped with the first descriptor loads.
//   iteration 0   every active propagator is evaluated once (store.rs:144-149 schedules
//                 all active propagators).  The per-family descriptor arrays are streamed
//                 from HBM by TMA bulk copies (cp.async.bulk + mbarrier, 4-stage ring in
//                 shared memory, one producer warp, 15 consumer warps of a 512-thread CTA
//                 with 128 registers per thread); domains are read from a shared-memory
//                 snapshot (V <= ~10k) or gathered from L2.  An all-XNeqY store streams a
//                 compact 8-byte copy of its descriptors.
//   iteration k   only propagators adjacent to variables that changed in iteration k-1
//                 are re-evaluated (store.rs:191-198 `react`): the dirty variables are a
//                 bit set that every CTA compacts into the same list; a short list is
//                 settled row by row (one CTA per variable, rounds out of shared memory,
//                 crawl shortcut), a long one is expanded warp by warp over the rows of
//                 the static var->propagator CSR (the reactor, reactors/indexed_deps.rs:
//                 23-27) with a per-propagator epoch stamp in the role of RelaxedFifo's
//                 `inside_queue` bit set (schedulers/relaxed_fifo.rs:42-48), and a very
//                 long one is swept again.
//
// `pcp_burst_kernel` wraps the same per-node code in a device-resident depth-first search
// (branching, label / restore on CTA 0; fast descent without a barrier).
//
// Updates are monotone atomics (atomicMax on lo, atomicMin on hi), so the chaotic
// iteration converges to the same greatest fixpoint as the reference FIFO (SURVEY 8a,
// "Parity theorem").  Integer bound arithmetic only: no tensor cores, no floating point.
#pragma once
#include "pcp_common.cuh"

// the full variant: every propagator family
namespace pcpd { namespace full {
#include "pcp_body.cuh"
} }  // namespace pcpd::full

// the binary-only variant (see pcp_body.cuh)
#define PCP_BIN_ONLY 1
namespace pcpd { namespace binonly {
#include "pcp_body.cuh"
} }  // namespace pcpd::binonly
#undef PCP_BIN_ONLY

// The same two variants for IntervalSet<i32> domains (PCP_FLAG_INTERVAL_SET; libpcp's VStoreSet,
// variable/mod.rs:38): a bit set per variable beside the cached bounds, XNeqY removes interior
// values (x_neq_y.rs:82-93 on IntervalSet::difference), bound updates land on the next value
// that is still in the set, entailment of XNeqY / XEqY / Distinct is decided on the sets.
// Kept as separate compilations so that the Interval kernels stay exactly what they were.
#define PCP_SET 1
namespace pcpd { namespace full_set {
#include "pcp_body.cuh"
} }  // namespace pcpd::full_set
#define PCP_BIN_ONLY 1
namespace pcpd { namespace binonly_set {
#include "pcp_body.cuh"
} }  // namespace pcpd::binonly_set
#undef PCP_BIN_ONLY
#undef PCP_SET
