/* pcp_b200.h -- C ABI of the B200-native propagation engine.
 *
 * This is the drop-in boundary for libpcp's hot path: everything at and below
 * `Consistency::consistency` (reference src/libpcp/kernel/consistency.rs:17-19,
 * implemented by src/libpcp/propagation/store.rs:240-258 and reached from
 * src/libpcp/search/space.rs:41-43) runs on the device behind these entry points.
 * The host search (libpcp::search) keeps driving branching and calls in here once
 * per search node.  INTEGRATION.md shows the Rust `extern "C"` block and the
 * `GpuCStore` shim that binds these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer argument is caller-owned and
 *     borrowed for the duration of the call; outputs go to caller buffers.
 *   - return value: PCP_OK or a negative error code.  A contract violation that
 *     panics in the reference (assert!) returns PCP_ERR_INVALID and leaves the
 *     engine unchanged.  Inconsistency is NOT an error: it is reported through
 *     *status == PCP_FALSE, exactly like `consistency() -> SKleene::False`.
 *   - one engine handle per host thread, bound to one CUDA device and one
 *     stream; a handle is not thread-safe, distinct handles are independent.
 *   - there is no CPU fallback: pcp_engine_create fails with PCP_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef PCP_B200_H
#define PCP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCP_OK 0
#define PCP_ERR_INVALID (-1)     /* contract violation (a panic in the reference)   */
#define PCP_ERR_CUDA (-2)        /* CUDA runtime / launch failure, or no device       */
#define PCP_ERR_NOMEM (-3)       /* host or device allocation failed                  */
#define PCP_ERR_UNSUPPORTED (-4) /* propagator/view that has no device lowering       */

/* trilean::SKleene as returned by Consistency::consistency. */
#define PCP_FALSE (-1)
#define PCP_UNKNOWN 0
#define PCP_TRUE 1

typedef struct pcp_engine pcp_engine;

/* A view over the variable store (reference src/libpcp/term/):
 *   var >= 0            Identity(var)            (term/identity.rs:47-70), off == 0
 *   var >= 0, off != 0  Addition(Identity, off)  (term/addition.rs:80-110; nested
 *                       additions fold into one offset)
 *   var == -1           Constant(off)            (term/constant.rs:43-68)
 *   var <= -2           Sum view #(-2 - var) + off  (term/sum.rs:56-92), see
 *                       pcp_sum_alloc.                                              */
typedef struct pcp_operand {
  int32_t var;
  int32_t off;
} pcp_operand;

#define PCP_VAR_CONSTANT (-1)
#define PCP_VAR_SUM(sum_id) (-2 - (sum_id))

/* Propagator kinds (reference src/libpcp/propagators/, src/libpcp/logic/). */
enum pcp_prop_kind {
  PCP_X_LESS_Y = 0,           /* cmp/x_less_y.rs:67-117; 2 operands (x, y)            */
  PCP_X_NEQ_Y = 1,            /* cmp/x_neq_y.rs:66-104; 2 operands                    */
  PCP_X_EQ_Y = 2,             /* cmp/x_eq_y.rs:67-116; 2 operands                     */
  PCP_X_GREATER_Y_PLUS_Z = 3, /* cmp/x_greater_y_plus_z.rs:75-128; 3 operands         */
  PCP_X_LESS_Y_PLUS_Z = 4,    /* cmp/x_less_y_plus_z.rs:75-128; 3 operands            */
  PCP_X_EQ_Y_PLUS_Z = 5,      /* cmp/x_eq_y_plus_z.rs:26-105; 3 operands              */
  PCP_DISTINCT = 6,           /* distinct.rs:69-126; n >= 1 operands                  */
  PCP_DISJ2_X_EQ_Y_PLUS_Z = 7,/* logic/disjunction.rs:77-129 over two XEqYPlusZ;
                                 6 operands (x1,y1,z1,x2,y2,z2)                       */
  PCP_X_EQ_Y_MUL_Z = 8,       /* cmp/x_eq_y_mul_z.rs:68-116; 3 operands: only x is narrowed,
                                 to x /\ (y * z) (interval product)                   */
  PCP_ALL_EQUAL = 9           /* all_equal.rs:47-103 (chain of XEqY); n >= 1 operands   */
};
#define PCP_NUM_KINDS 10

/* Engine configuration (the reference configures through type aliases only:
 * propagation/mod.rs:33-34, variable/mod.rs:35-38). */
#define PCP_FLAG_INCREMENTAL 1u /* skip the schedule-everything first sweep when the
                                   engine can prove the restored state was a fixpoint
                                   (same domains/status; fewer propagations than
                                   store.rs:144-149)                                  */
#define PCP_FLAG_INTERVAL_SET 4u /* domains are IntervalSet<i32> (VStoreSet, variable/mod.rs:38 -- what
                                   FDSpace and example/src/nqueens.rs use): a bit set per variable
                                   beside the cached bounds; XNeqY removes interior values.  Default:
                                   Interval<i32> (VStoreFD, variable/mod.rs:37)                   */
#define PCP_FLAG_HOST_SEARCH 2u /* pcp_search_* drive every node from the host through the
                                   store surface (restore / prop_alloc / consistency /
                                   domains_read / label, one launch and one D2H copy per
                                   node) -- the loop a libpcp host runs -- instead of the
                                   device-resident search                             */
typedef struct pcp_config {
  int32_t device;      /* CUDA device ordinal                                        */
  uint32_t flags;      /* PCP_FLAG_*                                                 */
  uint32_t max_labels; /* capacity of the label stack (0 = default 4096)             */
  uint32_t tail_limit; /* propagators allocated after the reactor CSR was built are kept in
                          a tail that every launch evaluates; the CSR is rebuilt when the
                          tail exceeds this (0 = default 4096)                        */
} pcp_config;

/* Counters the reference lacks (no propagation counter exists). */
typedef struct pcp_stats {
  uint64_t propagations; /* propagator evaluations (propagate + is_subsumed), this call */
  uint32_t iterations;   /* device fixpoint iterations (1 = first sweep was quiescent)  */
  uint32_t active_props; /* propagators still active (not entailed) after the call      */
  float kernel_ms;       /* device time of the fixpoint kernel (CUDA events); 0 unless
                            timing was enabled with pcp_set_timing                      */
  uint32_t launches;     /* persistent-kernel launches this call made for this engine (1; in a
                            batch served by one launch: 1 in stats[0], 0 in the others)    */
} pcp_stats;

int pcp_engine_create(const pcp_config* cfg, pcp_engine** out);
void pcp_engine_destroy(pcp_engine* e);
/* A second (vstore, cstore) pair over the same model: `Space::clone` without copying the model.  The
 * new engine starts from the parent's current state (domains, propagators, `active` set; no labels)
 * and shares the parent's immutable static part -- descriptors, reactor CSR -- on the device; tail,
 * domains, `active` set, trail and label stack are its own.  For the K subtree contexts of one GPU
 * (pcp_consistency_batch, pcp_search_step_many) this replaces K uploads and K reactor builds of the
 * same model by one, and keeps one copy of the descriptors in L2.  An engine that has to rebuild
 * its reactor (tail_limit reached) or is restored to a label older than the shared part lets go
 * of the shared block and uploads its own copy.  Call after the first pcp_consistency (the
 * reactor is built there); not while a search is open. */
int pcp_engine_fork(pcp_engine* parent, pcp_engine** out);
const char* pcp_last_error(const pcp_engine* e);
int pcp_set_timing(pcp_engine* e, int32_t enabled);
/* Several engines can run their fixpoints side by side on one GPU (independent subtrees of one
 * search: SURVEY 8e applied inside a device; pcp_consistency_batch).  A fixpoint launch is a
 * persistent grid of one CTA per SM; this caps the CTAs a launch of this engine uses so that the
 * engines' grids are co-resident (0 = all SMs, the default). */
int pcp_set_grid_limit(pcp_engine* e, int32_t max_ctas);
/* The CUDA stream (cudaStream_t, returned as an opaque pointer) every launch and copy of this
 * engine is issued on: lets a caller order its own device work -- e.g. a benchmark's L2
 * flush -- directly before a fixpoint without a host synchronisation in between. */
int pcp_stream(pcp_engine* e, void** stream);

/* VStore::alloc (variable/store.rs:135-140): n new variables with domains
 * [lo[i], hi[i]]; an empty domain is a contract violation (store.rs:136). */
int pcp_vars_alloc(pcp_engine* e, const int32_t* lo, const int32_t* hi, int32_t n, int32_t* first_idx);

/* Sum::new (term/sum.rs:28-32): registers a sum view over `n` (var, off) terms. */
int pcp_sum_alloc(pcp_engine* e, const pcp_operand* terms, int32_t n, int32_t* sum_id);

/* Store::alloc (propagation/store.rs:223-230): appends one propagator, marks it
 * active and returns its index. */
int pcp_prop_alloc(pcp_engine* e, int32_t kind, const pcp_operand* ops, int32_t n_ops, int32_t* idx);

/* ---- formula trees (logic/: Conjunction, Disjunction, Boolean, BooleanNeg, f.not()) ------------
 * Allocates ONE propagator whose body is a tree of connectives over leaf propagators, exactly as
 * `cstore.alloc(Box::new(Disjunction::new(vec![..])))` does in the reference (logic/disjunction.rs:
 * 29-33,77-129; logic/conjunction.rs:77-118; logic/boolean.rs:108-141; logic/boolean_neg.rs:77-98;
 * implication / equivalence = logic/mod.rs:30-45; the user of all of them: propagators/
 * cumulative.rs:59-114).  Prefix encoding in int32 words:
 *     PCP_F_CONJUNCTION n child_1 .. child_n          PCP_F_DISJUNCTION n child_1 .. child_n
 *     PCP_F_BOOLEAN var off                           PCP_F_BOOLEAN_NEG var off
 *     PCP_F_NOT child      (the reference's `f.not()`, applied when the tree is built: De Morgan on
 *                           connectives, the cmp/mod.rs:34-86 complements on leaves)
 *     PCP_F_LEAF + kind, then (var, off) per operand; kind one of PCP_X_LESS_Y, PCP_X_NEQ_Y,
 *                           PCP_X_EQ_Y, PCP_X_GREATER_Y_PLUS_Z, PCP_X_LESS_Y_PLUS_Z, PCP_X_EQ_Y_PLUS_Z
 * The operand of Boolean / BooleanNeg is a variable the caller allocated with domain [0, 1]
 * (Boolean::new, boolean.rs:35-40).  A tree whose subsumption / propagation the device cannot
 * evaluate (more than PCP_F_MAX_NODES nodes or PCP_F_MAX_VARS distinct variables, Sum operands,
 * not() of XEqYPlusZ -- unimplemented!() in the reference too) is PCP_ERR_UNSUPPORTED.           */
#define PCP_F_CONJUNCTION 1
#define PCP_F_DISJUNCTION 2
#define PCP_F_BOOLEAN 3
#define PCP_F_BOOLEAN_NEG 4
#define PCP_F_NOT 5
#define PCP_F_LEAF 16
#define PCP_F_MAX_NODES 32
#define PCP_F_MAX_VARS 12
int pcp_formula_alloc(pcp_engine* e, const int32_t* words, int32_t n_words, int32_t* idx);

/* The same for `n_props` propagators of one kind with `n_ops` operands each,
 * operands laid out prop-major (model upload: millions of descriptors). */
int pcp_props_alloc(pcp_engine* e, int32_t kind, const pcp_operand* ops, int32_t n_ops,
                    int64_t n_props, int32_t* first_idx);

/* Consistency::consistency (propagation/store.rs:247-257): runs the propagation
 * fixpoint on the device.  *status is PCP_FALSE / PCP_UNKNOWN / PCP_TRUE. */
int pcp_consistency(pcp_engine* e, int32_t* status, pcp_stats* stats /* may be NULL */);

/* Consistency::consistency for `n` engines at once: every engine's fixpoint is launched before the
 * first result is awaited, so the fixpoints of independent search nodes -- sibling subtrees held
 * by separate (vstore, cstore) pairs, the multi-GPU partitioning of SURVEY 8e applied inside one
 * GPU -- run side by side instead of one after the other (launch latency, device barriers and
 * worklist iterations of one node overlap the sweeps of the others).  status / stats have n
 * entries (stats may be NULL).  Engines without a pcp_set_grid_limit get an equal share of the
 * SMs for the call.  With timing enabled on engines[0], stats[0].kernel_ms is the device time of
 * the whole batch (first launch to last completion).
 * Engines that run the same kernel with the same geometry (forks of one model, snapshot of the
 * domains in shared memory, same grid limit, together no more CTAs than the device has SMs) are
 * served by ONE launch: consecutive groups of CTAs take one engine each.  K launches on K streams
 * start about 3 us apart on the device; the batch starts every engine's node at once. */
int pcp_consistency_batch(pcp_engine* const* engines, int32_t n, int32_t* status, pcp_stats* stats);

/* Index<usize> on the variable store (variable/store.rs:175-181), batched. */
int pcp_domains_read(pcp_engine* e, int32_t first, int32_t n, int32_t* lo, int32_t* hi);

/* Cardinality::size() of each domain (what FirstSmallestVar compares, search/branching/
 * first_smallest_var.rs:30-39): hi - lo + 1 on Interval domains, the number of values on
 * IntervalSet domains. */
int pcp_domains_size_read(pcp_engine* e, int32_t first, int32_t n, uint32_t* size);

/* The values of each domain as a bit window: bit (v - base) of the `words` 32-bit words of
 * variable i is set iff v is in its domain (IntervalSet iteration; on Interval engines one run
 * of ones).  Every value of the domains read must fall inside the window. */
int pcp_domains_read_bits(pcp_engine* e, int32_t first, int32_t n, int32_t base, int32_t words, uint32_t* out);

/* MonotonicUpdate::update (variable/store.rs:151-166): *ok = 0 when [lo,hi] is
 * empty (store untouched); widening is a contract violation. */
int pcp_var_update(pcp_engine* e, int32_t idx, int32_t lo, int32_t hi, int32_t* ok);

/* The `active` bit set of the constraint store (propagation/store.rs:34). */
int pcp_active_read(pcp_engine* e, int32_t first, int32_t n, uint8_t* out);

/* Snapshot::label / Snapshot::restore (kernel/restoration.rs:20-30) of the pair
 * (vstore, cstore) as in search/recomputation/no_recomputation.rs:49-61:
 * restore brings back the domains, truncates the propagators to the labelled
 * length and restores `active` (propagation/store.rs:312-323).  Labels are
 * plain values; restoring a label invalidates the labels taken after it. */
int pcp_label(pcp_engine* e, uint64_t* label);
int pcp_restore(pcp_engine* e, uint64_t label);

int pcp_num_vars(const pcp_engine* e, int32_t* n);
int pcp_num_props(const pcp_engine* e, int32_t* n);

/* ---- host search driver (callers of the path) ------------------
 * OneSolution/AllSolution o StopNode o [BranchAndBound o] Propagation o
 * Brancher(var_sel, val_sel, distributor) as in search/mod.rs:45-52, executed
 * node by node against the entry points above. */
typedef struct pcp_search_config {
  uint64_t node_limit;   /* StopNode (search/stop_node.rs:54-61); 0 = unlimited      */
  int32_t all_solutions; /* AllSolution (search/engine/all_solution.rs:41-50)        */
  int32_t var_sel;       /* 0 FirstSmallestVar, 1 InputOrder                          */
  int32_t val_sel;       /* 0 MiddleVal, 1 MinVal                                     */
  int32_t distributor;   /* 0 BinarySplit, 1 Enumerate                                */
  int32_t bb_mode;       /* 0 none, 1 minimize, 2 maximize (search/branch_and_bound.rs) */
  int32_t bb_var;
  int32_t trace_domains; /* 1: also copy lo/hi of every non-failed node into the trace */
  int32_t warmup_nodes;  /* the first warmup_nodes nodes are excluded from `seconds`,
                            `kernel_seconds`, `propagations` and `iterations`         */
} pcp_search_config;

typedef struct pcp_search_result {
  int32_t status;          /* 1 Satisfiable, -1 Unsatisfiable, 2 EndOfSearch          */
  int32_t has_bb_value;
  int32_t bb_value;
  int32_t reserved;
  uint64_t num_nodes, num_solution, num_failed_node, num_prune; /* search/statistics.rs:19-53 */
  uint64_t propagations;
  uint64_t iterations;
  double seconds;          /* host wall clock of the whole search                     */
  double kernel_seconds;   /* sum of device fixpoint kernel times (if timing enabled) */
} pcp_search_result;

/* Per-node trace (optional): status[i] in {-1,0,1}; hash[i] = FNV-1a over the
 * (lo,hi) words of all domains after the node's fixpoint (0 for failed nodes);
 * lo/hi (if trace_domains) hold num_vars entries per node. */
int pcp_search_run(pcp_engine* e, const pcp_search_config* cfg, pcp_search_result* res,
                   int32_t* trace_status, uint64_t* trace_hash, int32_t* trace_lo,
                   int32_t* trace_hi, uint64_t trace_capacity);

/* The same search as a resumable generator (OneSolution::enter can be called again,
 * search/engine/one_solution.rs:23): pcp_search_step runs at most `max_nodes` further
 * nodes (0 = no slice limit) and fills `res` with the cumulative counters;
 * res->status: 0 slice used up (search still open), 1 Satisfiable (one-solution mode;
 * step again for the next one), -1 tree exhausted without a further solution,
 * 2 EndOfSearch.  Between steps the host may exchange an incumbent / stop word with
 * other ranks.  The trace buffers must outlive the handle. */
typedef struct pcp_search pcp_search;
int pcp_search_open(pcp_engine* e, const pcp_search_config* cfg, int32_t* trace_status,
                    uint64_t* trace_hash, int32_t* trace_lo, int32_t* trace_hi,
                    uint64_t trace_capacity, pcp_search** out);
int pcp_search_step(pcp_search* s, uint64_t max_nodes, pcp_search_result* res);
/* pcp_search_step for `n` searches at once -- independent subtrees of one model, each opened on its
 * own engine (SURVEY 8e inside one GPU).  Device-resident searches whose engines run the same kernel
 * with the same geometry (forks of one model) share ONE launch per slice of nodes, a group of CTAs
 * each, every search running through its own budget at its own pace; otherwise one launch and one
 * host thread each.  Host-driven searches are shared out over PCP_SEARCH_THREADS host threads, each
 * keeping the next fixpoint of every search it drives in flight (PCP_SEARCH_LOCKSTEP=1: rounds of one
 * node per search through pcp_consistency_batch on the calling thread).  res has n entries (may be
 * NULL); a search that is finished or has used up `max_nodes` sits out the rest of the call. */
int pcp_search_step_many(pcp_search* const* searches, int32_t n, uint64_t max_nodes, pcp_search_result* res);
/* BranchAndBound's incumbent (search/branch_and_bound.rs:69-94) seen from outside: between two
 * pcp_search_step slices a rank adopts the best objective value any rank has found (the one-word
 * all-reduce of SURVEY 8e).  `value` replaces the local incumbent only when it improves on it
 * (smaller when minimising, larger when maximising); every node entered afterwards posts
 * `bb_var < value` (`>` when maximising) exactly as branch_and_bound.rs:76-87 does with its own
 * incumbent.  PCP_ERR_INVALID when the search was opened without bb_mode. */
int pcp_search_set_incumbent(pcp_search* s, int32_t value);
void pcp_search_close(pcp_search* s);

#ifdef __cplusplus
}
#endif
#endif /* PCP_B200_H */
