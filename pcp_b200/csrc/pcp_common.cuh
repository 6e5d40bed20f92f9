// pcp_common.cuh -- what both kernel variants share: launch parameters (Params / BurstParams),
// constants (ring geometry, family tags, descriptor encodings), the result header, the
// memory / mbarrier / TMA helpers and the two small utility kernels.  The kernels themselves
// are in pcp_body.cuh, compiled once per variant by pcp_device.cuh (which also carries the
// overview of the device side).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcpd {

constexpr int kThreads = 512;           // one CTA per SM, 128 registers per thread
constexpr int kWarps = kThreads / 32;
constexpr int kConsumerWarps = kWarps - 1;  // warp 0 drives the TMA ring
constexpr int kStages = 4;
constexpr int kStageBytes = 32768;
constexpr int kRingBytes = kStages * kStageBytes;
// propagators per ring stage: every consumer warp gets kGroupsBin / kGroupsTer whole groups of
// 32 consecutive propagators (= words of the `active` bit set) of each chunk
constexpr int kGroupsBin = 4, kGroupsTer = 2;
constexpr int kChunkBin = kConsumerWarps * 32 * kGroupsBin;  // 1920 * 16 B = 30720 B
constexpr int kChunkTer = kConsumerWarps * 32 * kGroupsTer;  // 960 * 16 B + 960 * 8 B
constexpr int kGroupsBinC = 8;           // compact 8-byte binary descriptors (Family::cdesc): twice the groups per stage
constexpr int kChunkBinC = kConsumerWarps * 32 * kGroupsBinC;  // 3840 * 8 B = 30720 B
constexpr int kSweepAllCtas = 32;       // groups up to this many CTAs: CTA 0 takes a share of the sweep too
constexpr int kChunkDj = 480;           // 480 * 48 B = 23040 B
static_assert(kChunkBin * 16 <= kStageBytes && kChunkTer * 16 <= 16384 && kChunkTer * 8 <= kStageBytes - 16384, "stage too small");
constexpr int kTerPlaneB = 16384;       // offset of the z plane inside a stage
constexpr unsigned kConstVar28 = 0x0FFFFFFFu;
constexpr unsigned kSumBase28 = 0x0F000000u;  // x-operand encodings >= this are sum views / Constant
constexpr int kMaxInline = 4;

// family tags inside adjacency / trail references (top 3 bits)
enum Fam : unsigned { F_BIN = 0, F_TER = 1, F_DJ = 2, F_NARY = 3 };
constexpr unsigned kSlotMask = 0x1FFFFFFFu;
__host__ __device__ inline unsigned make_ref(unsigned fam, unsigned slot) { return (fam << 29) | slot; }

// kinds inside a family (top 4 bits of descriptor word 0)
enum BinKind : unsigned { B_LESS = 0, B_NEQ = 1, B_EQ = 2 };
enum TerKind : unsigned { T_GREATER = 0, T_LESS = 1, T_EQ = 2, T_MUL = 3 };
enum NaryKind : int { N_DISTINCT = 0, N_ALL_EQUAL = 1, N_TREE = 2 };
// Formula trees (logic/: Conjunction, Disjunction, Boolean, BooleanNeg over cmp leaves) ride in the
// n-ary family: the operands of the n-ary slot are the tree's distinct variables, its nodes sit
// in Params::tree_nodes from tree_ptr[slot] on, kTreeNodeInts ints each, in prefix order:
//   [0] type: TN_* or TN_LEAF + BinKind / TN_LEAF3 + TerKind   [1] number of children
//   [2] index (relative to the tree's first node) one past the node's subtree
//   [3 + 2i], [4 + 2i]  operand i of a leaf / the Boolean: slot in the variable table (or -1:
//                       Constant) and offset
constexpr int kTreeNodeInts = 9;
constexpr int kTreeMaxNodes = 32, kTreeMaxVars = 12;
enum TreeNode : int { TN_CONJ = 1, TN_DISJ = 2, TN_BOOL = 3, TN_BOOL_NEG = 4, TN_LEAF = 16, TN_LEAF3 = 32 };

enum Decision : unsigned { D_CONTINUE = 0, D_FIXPOINT = 1, D_FAILED = 2, D_ITER_CAP = 3 };
// Worklist iterations: every CTA compacts the dirty bit set into a list that lives in the
// (idle) TMA ring; the rest of the ring stages the descriptors of a row for the row-local
// rounds.  A dirty set longer than the list is handled by a streaming sweep instead.
constexpr int kListCap = 8192;                       // dirty variables (32 KB of the ring)
constexpr int kRowStageOff = kListCap * 4;           // byte offset of the row staging area
constexpr unsigned kDecBits = 3;  // the release word of the barrier = (generation << kDecBits) | decision

struct Control {
  unsigned bar_count;
  unsigned bar_gen;
  unsigned decision[2];
  unsigned epoch;        // next unused epoch (stamps < epoch are stale)
  int failed;
  unsigned trail_cnt;    // number of deactivated (entailed) propagators on the trail
  unsigned trail_at_decision;  // trail_cnt as the last CTA to arrive at a deciding barrier saw it: the node's own
                               // entailments, all flushed, none of the next node's (the device search may run ahead)
  int pad0[2];
  unsigned long long propagations;  // cumulative
  unsigned iterations;   // of the last launch
  unsigned last_decision;
};

// Header copied back to the host together with the domains (one D2H copy per node).
struct Result {
  int failed;
  unsigned trail_cnt;
  unsigned iterations;
  unsigned epoch;
  unsigned long long propagations;
  unsigned decision;
  unsigned gen;          // barrier generation after the launch
  unsigned seq;          // host copy only: launch sequence number, written last (the host polls it)
  unsigned pad[7];
};
static_assert(sizeof(Result) == 64, "Result header is 64 bytes");

struct Family {
  int4* desc;           // BIN: 1 int4/prop; TER: int4 (x,y) plane; DJ: 3 int4/prop
  int2* descB;          // TER: (z) plane
  // The descriptors of the tail (slots >= n_static: branching constraints and whatever was posted
  // after the reactor CSR was built), indexed by slot like `desc`.  The same arrays as desc / descB
  // unless the static part is shared with other engines (pcp_engine_fork): then the tail is this
  // engine's own array, biased so that tdesc[slot * width] is slot's descriptor.
  int4* tdesc;
  int2* tdescB;
  uint32_t* active;     // bit set, 1 = active (propagation/store.rs:34)
  uint32_t* stamp;      // epoch of the last worklist evaluation
  int n;                // allocated propagators
  int n_static;         // [0, n_static) are covered by the CSR; [n_static, n) is the tail
  int all_plain;        // 1: no descriptor in [0, n_static) has a Constant or Sum operand
  int kind_mask;        // bit k: a descriptor of kind k may sit in [0, n_static) (over-approximation after restores)
  const uint2* cdesc;   // BIN: 8-byte copies {xvar | yvar << 16, (u16)xoff | yoff << 16} of [0, n_static) when all of
                        // them are XNeqY over plain variables with 16-bit ids / offsets, else nullptr
};

struct InlineProp {     // a propagator posted since the last launch, carried in the launch
  int4 q[3];            // parameters instead of a separate H2D copy
  unsigned fam;
  int slot;
  int pad[2];
};

struct Params {
  Result* result;
  int2* dom;            // interval per variable: (lo, hi)
  int V;
  int smem_dom;         // 1: domains are staged in shared memory
  int dirty_bm_off;     // byte offset (in dynamic shared memory) of the CTA's sweep dirty bitmap, 0 = none
  Family fam[3];        // BIN, TER, DJ
  const int* nary_ptr;  // CSR of n-ary Distinct operands
  const int2* nary_ops;
  const int* nary_kind;  // N_DISTINCT / N_ALL_EQUAL / N_TREE per n-ary propagator
  const int* tree_ptr;   // N_TREE: first int of the propagator's nodes in tree_nodes
  const int* tree_nodes;
  uint32_t* nary_active;
  int n_nary;
  int nary_max_k;
  const int* sum_ptr;   // CSR of the Sum views (term/sum.rs): terms are (var, off) operands
  const int2* sum_terms;
  const int* adj_ptr;   // reactor: var -> propagator refs
  const uint32_t* adj;
  uint32_t* dirty_bits; // 3 x dirty_words: bit v = variable v was narrowed (triple-buffered across iterations)
  int dirty_words;
  uint32_t* trail;
  Control* ctl;
  int full_sweep;       // 1: schedule every active propagator first (store.rs:144-149)
  unsigned max_iterations;
  unsigned gen0;        // generation of the device-wide barrier at launch (Result::gen of the last launch)
  unsigned epoch0;      // first unused stamp epoch (Result::epoch of the last launch)
  // ---- node prologue (Snapshot::restore + Store::alloc), executed by CTA 0
  const int2* restore_from;   // label copy of the domains, or nullptr
  unsigned trail_keep;        // trail length recorded in the label
  int do_trail;               // 1: re-activate trail entries >= trail_keep (store.rs:319-323)
  int sync0;                  // 1: the prologue has cross-CTA effects -> barrier before use
  uint32_t* nary_active_w;    // == nary_active (writable alias for the prologue)
  int new_first[4], new_last[4];  // slots whose active bit must be set (per family)
  int n_inline;
  InlineProp inl[kMaxInline];
  int seed_dirty;             // incremental launch: dirty_bits[0] seeded by the host (count, may be 0)
  // ---- epilogue
  int2* snapshot_to;          // if not failed: copy of the fixpoint domains (next label slot)
  Result* host_result;        // mapped pinned host memory: result header (+ domains behind it when
  int host_dom;               //   host_dom = 1), stored by the kernel itself and published through
  unsigned host_seq;          //   Result::seq, so the host polls instead of copying + synchronising
  // ---- debug: per-CTA phase timestamps (PCP_TRACE=1), 8 slots per CTA
  unsigned long long* trace;
  // ---- IntervalSet domains (PCP_FLAG_INTERVAL_SET; the `set` kernel variants): one bit per value
  // of the window [bits_base, bits_base + 32 * bits_W) and variable; the domain of v is
  // bits[v] /\ [dom[v].lo, dom[v].hi] (bits outside the cached bounds are don't-care).  A label
  // slot holds the V domains followed by the V * bits_W words (stack_stride counts both).
  uint32_t* bits;
  int bits_W;
  int bits_base;
  uint32_t* host_sizes;       // mapped pinned: Cardinality::size() per variable, stored by the epilogue
};

constexpr int kTraceIter1 = 8 * 256 + 4 * 32 + 64;  // second mark region: phases of iteration 1
constexpr int kTraceSeq = kTraceIter1 + 8 * 256;   // CTA 0, iteration 1: (id, time) sequence inside the row-local rounds
constexpr int kTraceWords = kTraceSeq + 2 + 2 * 64;
__device__ __forceinline__ void trace_mark(const Params& P, int slot) {
  if (P.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
    P.trace[blockIdx.x * 8 + slot] = t;
  }
}
__device__ __forceinline__ void trace_seq(const Params& P, unsigned id, unsigned dep = 0u) {
  if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && P.trace[kTraceSeq + 1] == 1ull) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer; // %1" : "=l"(t) : "r"(dep) : "memory");  // `dep`: wait for that register
    const unsigned long long n = P.trace[kTraceSeq];
    if (n < 64) { P.trace[kTraceSeq + 2 + 2 * n] = id; P.trace[kTraceSeq + 3 + 2 * n] = t; P.trace[kTraceSeq] = n + 1; }
  }
}
__device__ __forceinline__ void trace_mark1(const Params& P, unsigned iter, int slot) {
  if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && slot == 0) P.trace[kTraceSeq + 1] = iter == 1 ? 1ull : 0ull;
  if (P.trace && iter == 1 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
    P.trace[kTraceIter1 + blockIdx.x * 8 + slot] = t;
  }
}

// ---------------------------------------------------------------------------------------
// memory / async helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int2 ldcg_dom(const int2* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}


struct BurstCtl {
  int cmd;            // 0: run the posted node, 1: stop
  int root_pending;   // 1: the next node is the root of the search (nothing to pop)
  int inl_slot;       // slot of the posted branching constraint (-1: none = root)
  int bin_n;          // end of the binary tail for the posted node
  int4 inl_desc;
  int n_branch, n_labels, cur_label, stopped;
  int status;         // why the burst ended: 0 budget, 1 solution, -1 exhausted, 2 end of search
  int err;            // 1: label / branch / tail capacity exceeded
  int last_status;    // status of the last node (-1/0/1)
  int pad;
  unsigned long long nodes, solutions, failures, iterations;
};

struct BurstParams {
  BurstCtl* bc;
  int4* branches;          // DFS stack: (label, var, val, alternative)
  int2* label_meta;        // per label: (bin_n, trail_len)
  int2* branch_meta;       // the same record next to every branch (one round trip per pop)
  int2* stack;             // label slots (copies of dom[])
  long long stack_stride;
  int max_labels, max_branches, bin_cap;
  int all_solutions;
  int incremental;                  // PCP_FLAG_INCREMENTAL: nodes below the root evaluate their posted constraint and what it wakes
  unsigned long long node_budget;   // nodes to run in this launch
  unsigned long long node_limit;    // StopNode (search/stop_node.rs:54-61), 0 = none
  long long props_base;             // allocated propagators = props_base + bin_n
  int* t_status;                    // per-node trace (device buffers), indexed by node number
  int2* t_dom;
  uint32_t* t_bits;                 // set variants: V * bits_W words per traced node
  unsigned long long t_cap;
};

// CTA 0's working copy of the search state: lives in shared memory for the whole burst so
// that the host work between two nodes costs one or two global round trips.
struct BurstLocal {
  int n_branch, n_labels, cur_label, bin_n, stopped, root_pending, last_status, err;
  unsigned long long nodes, solutions, failures, iterations;
  int4 top;        // the branch pushed last (the left child) and its label record
  int2 top_meta;
  int top_valid;
  int run;         // set by burst_host_step: 1 = a node was posted
};

__global__ void pcp_fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// incremental launches: the variables narrowed by the host enter the dirty set of iteration 0
__global__ void pcp_seed_dirty_kernel(const int* list, int n, uint32_t* bits0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicOr(&bits0[list[i] >> 5], 1u << (list[i] & 31));
}

}  // namespace pcpd
