// pcp_engine.cu -- host side of the B200 propagation engine and its C ABI
// (include/pcp_b200.h).  The engine handle owns every device allocation; the reference's
// constraint store (src/libpcp/propagation/store.rs:32-37) and variable store
// (src/libpcp/variable/store.rs:29-33) map to:
//
//   domains        int2 dom[V] = (lo, hi) per variable, in HBM, preceded by a 64-byte
//                  result header so that status + domains come back in one D2H copy
//   propagators    per-family descriptor arrays (binary 16 B, ternary 16 B + 8 B planes,
//                  2-way disjunction 48 B, n-ary Distinct as CSR of operands)
//   active         one bit per propagator and family (store.rs:34) + a trail of the
//                  propagators entailed so far (restored by label, store.rs:319-323)
//   reactor        static CSR var -> propagator refs (reactors/indexed_deps.rs:23-27),
//                  built once per model; propagators allocated later form a short tail
//   labels         copy stack of the domain array (variable/memory/copy_memory.rs:141-151)
//
// A search node costs one kernel launch and one D2H copy: restore, the posted branching
// constraint and the label copy all ride inside the fixpoint launch (pcp_device.cuh).
// There is no CPU path: without a usable CUDA device pcp_engine_create fails.
#include "../../include/pcp_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "pcp_device.cuh"
#include "pcp_internal.h"

using namespace pcpd;

namespace {

struct Error {
  int code;
  std::string msg;
};
#define PCP_FAIL(code, msg) throw Error{code, msg}
#define PCP_REQUIRE(cond, msg) \
  do { if (!(cond)) PCP_FAIL(PCP_ERR_INVALID, msg); } while (0)
// While a device-resident search is open its launch parameters cache raw device pointers and the
// search state lives on the device: the store entry points that would reallocate or change
// either are contract violations until pcp_search_close.
#define PCP_REQUIRE_NO_BURST(e) PCP_REQUIRE(!(e)->burst.open, "a device search is open on this engine (pcp_search_close it first)")
#define CUDA_CHECK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      PCP_FAIL(_e == cudaErrorMemoryAllocation ? PCP_ERR_NOMEM : PCP_ERR_CUDA,                \
               std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
  } while (0)

constexpr size_t kPad = 64;  // slack elements behind every device array (TMA tails)
// All device bound arithmetic (view offsets, y.hi - 1, y.lo + z.lo + strict, the segmented sum
// of a Sum view, MiddleVal's lo + hi) is plain 32-bit: every bound, offset and worst-case Sum
// range must stay below 2^29 in magnitude.  Rust release builds wrap and debug builds panic
// on overflow (SURVEY 7 "hard parts" iv), so such inputs are outside the pinned contract:
// they are rejected with PCP_ERR_INVALID instead of wrapping silently.
constexpr long long kBoundLimit = 1ll << 29;

// Growable device array.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  void reserve(size_t n, cudaStream_t st, size_t keep_n = 0) {
    if (n <= cap) return;
    size_t ncap = std::max<size_t>(n, cap + cap / 2 + 64);
    T* q = nullptr;
    if (std::getenv("PCP_DEBUG_ALLOC")) std::fprintf(stderr, "[pcp alloc] %zu -> %zu elements of %zu bytes\n", cap, ncap, sizeof(T));
    CUDA_CHECK(cudaMalloc(&q, (ncap + kPad) * sizeof(T)));
    try {
      CUDA_CHECK(cudaMemsetAsync(q + ncap, 0, kPad * sizeof(T), st));
      if (p) {
        if (keep_n) CUDA_CHECK(cudaMemcpyAsync(q, p, keep_n * sizeof(T), cudaMemcpyDeviceToDevice, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
      }
    } catch (...) {
      cudaFree(q);  // the old array stays valid
      throw;
    }
    if (p) cudaFree(p);
    p = q;
    cap = ncap;
  }
  void free() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostFamily {
  std::vector<int4> desc;   // BIN: n, TER: n, DJ: 3n
  std::vector<int2> descB;  // TER only
  size_t n = 0;
  size_t n_static = 0;      // covered by the CSR
  size_t uploaded = 0;      // descriptors [0, uploaded) are on the device
  size_t active_set = 0;    // active bits [0, active_set) are initialised on the device
  size_t first_nonplain = SIZE_MAX;  // first descriptor with a Constant / Sum operand (lean sweep variant)
  int kind_mask = 0;                 // kinds ever allocated in this family
  int static_kind_mask = 0;          // ... among the descriptors covered by the CSR (over-approximation)
  DevBuf<int4> d_desc;      // static part + tail; a *view* into StaticBlock when the engine shares its static model
  DevBuf<int2> d_descB;
  DevBuf<int4> d_tdesc;     // shared static model only: this engine's tail, slot s at [(s - tail_base) * width]
  DevBuf<int2> d_tdescB;
  size_t tail_base = 0;     // first slot of d_tdesc (= n_static when the static part became shared)
  DevBuf<uint2> d_cdesc;    // BIN only: compact 8-byte copies of the static descriptors (see build_csr)
  size_t n_cdesc = 0;       // descriptors covered by d_cdesc (0: none / not compactable)
  DevBuf<uint32_t> d_active, d_stamp;
  int width = 1;
};

// The immutable part of a loaded model -- static descriptors, compact stream, reactor CSR -- once it
// is shared between engines (pcp_engine_fork): freed with the last engine that refers to it.
struct StaticBlock {
  DevBuf<int4> desc[3];
  DevBuf<int2> descB;
  DevBuf<uint2> cdesc;
  DevBuf<int> adj_ptr;
  DevBuf<uint32_t> adj;
  ~StaticBlock() {
    for (auto& d : desc) d.free();
    descB.free(); cdesc.free(); adj_ptr.free(); adj.free();
  }
};

struct LabelRec {
  size_t n_fam[4];
  size_t n_nary_ops;
  size_t n_tree_ints;
  size_t n_props;
  unsigned trail_len;
  bool at_fixpoint;
  uint64_t dom_version;     // version of the domains captured by this label
};

}  // namespace

struct pcp_engine {
  int device = 0;
  uint32_t flags = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_sleep = nullptr;   // cudaEventBlockingSync: waits that put the host thread to sleep
  cudaEvent_t ev_done = nullptr, ev_batch0 = nullptr, ev_batch1 = nullptr;  // pcp_consistency_batch: fork / join across engines' streams
  Params* d_batch = nullptr;  // pcp_consistency_batch, one launch for several engines: their launch parameters
  Params* h_batch = nullptr;  // (pinned staging of the same)
  int batch_cap = 0;
  void* d_bbatch = nullptr;   // pcp_search_step_many, one launch for several device searches: parameters + results
  void* h_bbatch = nullptr;
  int burst_batch_cap = 0;
  bool timing = false;
  int num_sms = 0;
  int grid_limit = 0;               // pcp_set_grid_limit: CTAs a launch of this engine may use (0 = all SMs)
  int max_smem_optin = 0;
  size_t static_smem = 0;           // largest static shared memory of the persistent kernels
  std::string err;

  // variables
  std::vector<int2> h_dom_pending;  // variables allocated but not yet uploaded
  std::vector<int2> h_dom_init;     // the domains as allocated (range validation of Sum views)
  size_t V = 0, V_uploaded = 0;
  // IntervalSet domains (PCP_FLAG_INTERVAL_SET): one bit per value of the window
  // [set_base, set_base + 32 * set_W) and variable, fixed when the first variables are uploaded
  bool set_mode = false;
  int set_base = 0, set_W = 0;
  DevBuf<uint32_t> d_bits;
  bool sizes_valid = false;         // the size mirror behind the host domains is current
  char* d_block = nullptr;          // [Result | dom ...]
  size_t block_cap_vars = 0;
  char* h_block = nullptr;          // pinned mirror of d_block (mapped: the fixpoint kernel stores into it)
  char* h_block_dev = nullptr;      // its device-side address
  unsigned host_seq = 0;            // sequence number of the last zero-copy result
  size_t h_block_cap_vars = 0;
  bool mirror_valid = false;        // h_block domains == device domains
  uint64_t dom_version = 1;         // bumped whenever the device domains may change

  // propagators
  HostFamily fam[3];                // BIN, TER, DJ
  std::vector<int> h_nary_ptr{0};
  std::vector<int2> h_nary_ops;
  std::vector<int> h_nary_kind;     // per n-ary propagator: N_DISTINCT / N_ALL_EQUAL / N_TREE
  DevBuf<int> d_nary_kind;
  std::vector<int> h_tree_ptr;      // per n-ary propagator: first int of its nodes in h_tree_nodes (-1: not a tree)
  std::vector<int> h_tree_nodes;    // formula trees, kTreeNodeInts ints per node (pcp_common.cuh)
  DevBuf<int> d_tree_ptr, d_tree_nodes;
  size_t n_nary = 0, nary_uploaded = 0, nary_active_set = 0;
  int nary_max_k = 0;
  DevBuf<int> d_nary_ptr;
  DevBuf<int2> d_nary_ops;
  DevBuf<uint32_t> d_nary_active;
  std::vector<uint32_t> prop_ref;   // global propagator index -> (family, slot)
  std::vector<std::vector<pcp_operand>> sums;

  std::vector<int> h_sum_ptr{0};
  std::vector<int2> h_sum_terms;
  DevBuf<int> d_sum_ptr;
  DevBuf<int2> d_sum_terms;
  size_t sums_uploaded = 0;

  // reactor CSR
  DevBuf<int> d_adj_ptr;            // (views into `shared` when set)
  DevBuf<uint32_t> d_adj;
  std::shared_ptr<StaticBlock> shared;  // the static model is shared with other engines (pcp_engine_fork)
  bool csr_built = false;
  size_t tail_limit = 4096;

  // worklists / control
  DevBuf<uint32_t> d_dirty_bits;  // 3 x dirty_words, all zero between launches
  size_t dirty_words = 0;
  DevBuf<int> d_seed_list;        // host-narrowed variables of an incremental launch
  DevBuf<uint32_t> d_trail;
  Control* d_ctl = nullptr;
  unsigned trail_len = 0;           // host view of ctl->trail_cnt
  unsigned epoch = 1;
  unsigned bar_gen = 0;             // generation of the device-wide barrier (Result::gen of the last launch)
  unsigned long long props_total = 0;

  // labels
  std::vector<LabelRec> labels;
  DevBuf<int2> d_stack;             // label slots, stride = stack_stride
  size_t stack_stride = 0;
  size_t max_labels = 4096;
  bool snapshot_valid = false;      // slot labels.size() holds the domains of dom_version
  uint64_t snapshot_version = 0;

  // deferred node prologue
  bool pending_restore = false;
  size_t pending_restore_label = 0;
  bool pending_trail_undo = false;
  unsigned pending_trail_keep = 0;
  bool at_fixpoint = false;
  std::vector<int> host_dirty;      // variables narrowed through pcp_var_update
  unsigned max_iterations = 1u << 22;

  // device-resident search bursts (pcp_internal.h)
  struct Burst {
    bool open = false;
    uint64_t root_label = 0;
    BurstCtl* d_bc = nullptr;
    DevBuf<int4> d_branches;
    DevBuf<int2> d_meta, d_bmeta;
    DevBuf<int> d_tstatus;
    DevBuf<int2> d_tdom;
    DevBuf<uint32_t> d_tbits;       // IntervalSet engines: the bit sets of the traced nodes
    uint64_t trace_cap = 0;
    bool trace_dom = false;
    int all_solutions = 0;
    uint64_t node_limit = 0;
    int max_labels = 0, bin_cap = 0;
    long long props_base = 0;
    unsigned long long props0 = 0;
    double kernel_seconds = 0;
    Params P;
  } burst;

  // the fixpoint launched by fixpoint_launch and not yet collected by fixpoint_wait
  struct Inflight {
    bool prepared = false, active = false;
    Params P;
    const void* fn = nullptr;
    size_t smem = 0;
    int grid = 0;
    bool want_snapshot = false, eager_dom = false, zero_copy = false;
    bool fused = false;  // launched as one group of a batched launch (no events of its own)
    double hp0 = 0, hp1 = 0, hp2 = 0, hp3 = 0;
  } inflight;

  // debug timeline (PCP_TRACE=1)
  unsigned long long* d_trace = nullptr;

  // pinned staging for small uploads
  char* h_stage = nullptr;
  size_t h_stage_cap = 0;

  Result* d_result() const { return reinterpret_cast<Result*>(d_block); }
  int2* d_dom() const { return reinterpret_cast<int2*>(d_block + sizeof(Result)); }
  Result* h_result() const { return reinterpret_cast<Result*>(h_block); }
  int2* h_dom() const { return reinterpret_cast<int2*>(h_block + sizeof(Result)); }
  uint32_t* h_sizes() const { return reinterpret_cast<uint32_t*>(h_block + sizeof(Result) + h_block_cap_vars * sizeof(int2)); }
  uint32_t* h_sizes_dev() const { return reinterpret_cast<uint32_t*>(h_block_dev + sizeof(Result) + h_block_cap_vars * sizeof(int2)); }
  // a label slot: the V domains, then (set mode) the V * set_W words of the bit sets, in int2 units
  size_t slot_stride() const { return V + (set_mode ? (V * (size_t)set_W + 1) / 2 : 0); }
  size_t num_props() const { return prop_ref.size(); }
};

namespace {

void ensure_var_capacity(pcp_engine* e, size_t nvars) {
  if (nvars > e->block_cap_vars) {
    size_t ncap = std::max<size_t>(nvars, e->block_cap_vars * 2 + 256);
    char* q = nullptr;
    CUDA_CHECK(cudaMalloc(&q, sizeof(Result) + ncap * sizeof(int2)));
    CUDA_CHECK(cudaMemsetAsync(q, 0, sizeof(Result), e->stream));
    if (e->d_block) {
      if (e->V_uploaded)
        CUDA_CHECK(cudaMemcpyAsync(q + sizeof(Result), e->d_block + sizeof(Result), e->V_uploaded * sizeof(int2),
                                   cudaMemcpyDeviceToDevice, e->stream));
      CUDA_CHECK(cudaStreamSynchronize(e->stream));
      cudaFree(e->d_block);
    }
    e->d_block = q;
    e->block_cap_vars = ncap;
  }
  if (nvars > e->h_block_cap_vars) {
    size_t ncap = std::max<size_t>(nvars, e->h_block_cap_vars * 2 + 256);
    char* q = nullptr;
    CUDA_CHECK(cudaHostAlloc(&q, sizeof(Result) + ncap * (sizeof(int2) + sizeof(uint32_t)), cudaHostAllocMapped));
    e->sizes_valid = false;
    if (e->h_block) {
      std::memcpy(q, e->h_block, sizeof(Result) + e->h_block_cap_vars * sizeof(int2));
      cudaFreeHost(e->h_block);
    } else {
      std::memset(q, 0, sizeof(Result));
    }
    e->h_block = q;
    e->h_block_cap_vars = ncap;
    void* dp = nullptr;
    CUDA_CHECK(cudaHostGetDevicePointer(&dp, q, 0));
    e->h_block_dev = static_cast<char*>(dp);
  }
}

void* stage(pcp_engine* e, size_t bytes) {
  if (bytes > e->h_stage_cap) {
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    if (e->h_stage) cudaFreeHost(e->h_stage);
    size_t ncap = std::max<size_t>(bytes, 1 << 16);
    CUDA_CHECK(cudaMallocHost(&e->h_stage, ncap));
    e->h_stage_cap = ncap;
  }
  return e->h_stage;
}

// Upload host range [from, to) of a vector to the device array.
template <class T>
void upload_range(pcp_engine* e, DevBuf<T>& dst, const std::vector<T>& src, size_t from, size_t to) {
  if (to <= from) return;
  if (src.size() > dst.cap) dst.reserve(src.size() + 3 * e->tail_limit, e->stream, from);  // (headroom: see prepare)
  CUDA_CHECK(cudaMemcpyAsync(dst.p + from, src.data() + from, (to - from) * sizeof(T), cudaMemcpyHostToDevice, e->stream));
  CUDA_CHECK(cudaStreamSynchronize(e->stream));  // pageable source: keep it simple and safe
}

void fill_u32(pcp_engine* e, uint32_t* p, uint32_t v, size_t n) {
  if (!n) return;
  int blocks = (int)std::min<size_t>((n + 255) / 256, 2048);
  pcp_fill_u32_kernel<<<blocks, 256, 0, e->stream>>>(p, v, n);
  CUDA_CHECK(cudaGetLastError());
}

// Grow a per-propagator bit set / stamp array, zero-filling the new part.
void reserve_zeroed(pcp_engine* e, DevBuf<uint32_t>& b, size_t words, size_t valid_words) {
  if (words <= b.cap) return;
  size_t from = std::min(valid_words, b.cap);
  b.reserve(words, e->stream, from);
  fill_u32(e, b.p + from, 0u, b.cap - from);
}

inline unsigned enc_var28(int var) {
  if (var >= 0) return (unsigned)var;
  return var == -1 ? kConstVar28 : kSumBase28 + (unsigned)(-2 - var);
}

void check_operand(const pcp_engine* e, pcp_operand op) {
  PCP_REQUIRE(op.off > -kBoundLimit && op.off < kBoundLimit, "view offset / constant outside (-2^29, 2^29)");
  if (op.var >= 0) {
    PCP_REQUIRE((size_t)op.var < e->V, "operand variable not registered in the store");
    const int2 d = e->h_dom_init[(size_t)op.var];  // the view's range (domains only shrink)
    PCP_REQUIRE((long long)d.x + op.off > -kBoundLimit && (long long)d.y + op.off < kBoundLimit,
                "range of the view var + offset outside (-2^29, 2^29)");
  } else if (op.var <= -2) PCP_REQUIRE((size_t)(-2 - op.var) < e->sums.size(), "unknown sum view");
  PCP_REQUIRE(op.var < (int)kSumBase28, "variable index too large");
}

// Sum views with a single term delegate to the term (term/sum.rs:62-64); wider sums stay
// read-only views (var = -2 - sum id) that the device evaluates by a segmented sum.
pcp_operand lower_view(const pcp_engine* e, pcp_operand op) {
  check_operand(e, op);
  if (op.var > -2) return op;
  const auto& s = e->sums[(size_t)(-2 - op.var)];
  if (s.size() == 1) {
    pcp_operand t = s[0];
    t.off += op.off;
    return t;
  }
  PCP_REQUIRE((size_t)(-2 - op.var) < (size_t)(kConstVar28 - kSumBase28), "too many sum views");
  return op;
}

// The variables a lowered operand depends on (ViewDependencies, term/ops.rs:26-28).
template <class F>
void for_each_dep_var(const pcp_engine* e, int var, F&& fn) {
  if (var >= 0) fn(var);
  else if (var <= -2)
    for (const pcp_operand& t : e->sums[(size_t)(-2 - var)])
      if (t.var >= 0) fn(t.var);
}

void require_distinct_vars(const pcp_engine* e, const pcp_operand* ops, int n) {
  // a propagator subscribing twice to one variable panics in the reference
  // (reactors/indexed_deps.rs:69-77)
  bool any_sum = false;
  for (int i = 0; i < n; ++i) any_sum |= ops[i].var <= -2;
  if (!any_sum) {  // the bulk-upload path: no allocation
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j)
        PCP_REQUIRE(ops[i].var < 0 || ops[i].var != ops[j].var, "propagator already subscribed to this variable");
    return;
  }
  std::vector<int> vs;
  for (int i = 0; i < n; ++i) for_each_dep_var(e, ops[i].var, [&](int v) { vs.push_back(v); });
  std::sort(vs.begin(), vs.end());
  PCP_REQUIRE(std::adjacent_find(vs.begin(), vs.end()) == vs.end(), "propagator already subscribed to this variable");
}

void append_prop(pcp_engine* e, int kind, const pcp_operand* raw, int n_ops) {
  pcp_operand ops[6];
  if (e->set_mode) {
    // What has no device lowering on IntervalSet domains (loud, not a fallback): the arithmetic of
    // XEqYMulZ and of multi-term Sum views needs `set * set` / `set + set` (whose reference result
    // depends on interior values while the propagators subscribe to Bound only: the reference's
    // own fixpoint is schedule-dependent there), AllEqual needs an n-way set intersection.
    if (kind == PCP_X_EQ_Y_MUL_Z || kind == PCP_ALL_EQUAL)
      PCP_FAIL(PCP_ERR_UNSUPPORTED, "XEqYMulZ / AllEqual have no device lowering on IntervalSet domains");
    for (int i = 0; i < n_ops; ++i)
      if (raw[i].var <= -2 && (size_t)(-2 - raw[i].var) < e->sums.size() && e->sums[(size_t)(-2 - raw[i].var)].size() > 1)
        PCP_FAIL(PCP_ERR_UNSUPPORTED, "multi-term Sum views have no device lowering on IntervalSet domains");
  }
  switch (kind) {
    case PCP_X_LESS_Y:
    case PCP_X_NEQ_Y:
    case PCP_X_EQ_Y: {
      PCP_REQUIRE(n_ops == 2, "binary propagator takes 2 operands");
      for (int i = 0; i < 2; ++i) ops[i] = lower_view(e, raw[i]);
      require_distinct_vars(e, ops, 2);
      unsigned k = kind == PCP_X_LESS_Y ? B_LESS : (kind == PCP_X_NEQ_Y ? B_NEQ : B_EQ);
      HostFamily& f = e->fam[F_BIN];
      if (ops[0].var < 0 || ops[1].var < 0) f.first_nonplain = std::min(f.first_nonplain, f.n);
      f.kind_mask |= 1 << k;
      f.desc.push_back(make_int4((int)((k << 28) | enc_var28(ops[0].var)), ops[0].off, ops[1].var, ops[1].off));
      e->prop_ref.push_back(make_ref(F_BIN, (unsigned)f.n++));
      break;
    }
    case PCP_X_GREATER_Y_PLUS_Z:
    case PCP_X_LESS_Y_PLUS_Z:
    case PCP_X_EQ_Y_PLUS_Z:
    case PCP_X_EQ_Y_MUL_Z: {
      PCP_REQUIRE(n_ops == 3, "ternary propagator takes 3 operands");
      for (int i = 0; i < 3; ++i) ops[i] = lower_view(e, raw[i]);
      require_distinct_vars(e, ops, 3);
      unsigned k = kind == PCP_X_GREATER_Y_PLUS_Z ? T_GREATER
                   : (kind == PCP_X_LESS_Y_PLUS_Z ? T_LESS : (kind == PCP_X_EQ_Y_PLUS_Z ? T_EQ : T_MUL));
      HostFamily& f = e->fam[F_TER];
      if (ops[0].var < 0 || ops[1].var < 0 || ops[2].var < 0) f.first_nonplain = std::min(f.first_nonplain, f.n);
      f.kind_mask |= 1 << k;
      f.desc.push_back(make_int4((int)((k << 28) | enc_var28(ops[0].var)), ops[0].off, ops[1].var, ops[1].off));
      f.descB.push_back(make_int2(ops[2].var, ops[2].off));
      e->prop_ref.push_back(make_ref(F_TER, (unsigned)f.n++));
      break;
    }
    case PCP_DISJ2_X_EQ_Y_PLUS_Z: {
      PCP_REQUIRE(n_ops == 6, "2-way disjunction of XEqYPlusZ takes 6 operands");
      for (int i = 0; i < 6; ++i) ops[i] = lower_view(e, raw[i]);
      require_distinct_vars(e, ops, 3);
      require_distinct_vars(e, ops + 3, 3);
      HostFamily& f = e->fam[F_DJ];
      f.desc.push_back(make_int4(ops[0].var, ops[0].off, ops[1].var, ops[1].off));
      f.desc.push_back(make_int4(ops[2].var, ops[2].off, ops[3].var, ops[3].off));
      f.desc.push_back(make_int4(ops[4].var, ops[4].off, ops[5].var, ops[5].off));
      e->prop_ref.push_back(make_ref(F_DJ, (unsigned)f.n++));
      break;
    }
    case PCP_DISTINCT:
    case PCP_ALL_EQUAL: {
      PCP_REQUIRE(n_ops >= 1, kind == PCP_DISTINCT ? "Variable array in `Distinct` must be non-empty."
                                                   : "Variable array in `AllEqual` must be non-empty.");
      PCP_REQUIRE(n_ops <= 4096, "n-ary propagator over more than 4096 operands is not supported");
      std::vector<pcp_operand> lo(n_ops);
      for (int i = 0; i < n_ops; ++i) {
        lo[i] = lower_view(e, raw[i]);
        if (lo[i].var <= -2) PCP_FAIL(PCP_ERR_UNSUPPORTED, "n-ary propagators over multi-term Sum views have no device lowering");
      }
      {
        std::vector<int> vs;
        for (auto& o : lo) if (o.var >= 0) vs.push_back(o.var);
        std::sort(vs.begin(), vs.end());
        PCP_REQUIRE(std::adjacent_find(vs.begin(), vs.end()) == vs.end(), "propagator already subscribed to this variable");
      }
      for (auto& o : lo) e->h_nary_ops.push_back(make_int2(o.var, o.off));
      e->h_nary_ptr.push_back((int)e->h_nary_ops.size());
      e->h_nary_kind.push_back(kind == PCP_DISTINCT ? (int)N_DISTINCT : (int)N_ALL_EQUAL);
      e->h_tree_ptr.push_back(-1);
      e->nary_max_k = std::max(e->nary_max_k, n_ops);
      e->prop_ref.push_back(make_ref(F_NARY, (unsigned)e->n_nary++));
      // binary/ternary/disjunction propagators allocated after a fixpoint land in the tail,
      // which every launch evaluates; a new n-ary propagator needs a full first sweep
      e->at_fixpoint = false;
      break;
    }
    default:
      PCP_FAIL(PCP_ERR_UNSUPPORTED, "unknown propagator kind");
  }
}

void unshare_static(pcp_engine* e);
void truncate_props(pcp_engine* e, const LabelRec& r) {
  if (e->shared)
    for (int f = 0; f < 3; ++f)
      if (r.n_fam[f] < e->fam[f].tail_base) { unshare_static(e); break; }
  for (int f = 0; f < 3; ++f) {
    HostFamily& hf = e->fam[f];
    hf.n = r.n_fam[f];
    hf.desc.resize(hf.n * hf.width);
    if (f == F_TER) hf.descB.resize(hf.n);
    hf.n_static = std::min(hf.n_static, hf.n);
    hf.uploaded = std::min(hf.uploaded, hf.n);
    hf.active_set = std::min(hf.active_set, hf.n);
    if (hf.first_nonplain >= hf.n) hf.first_nonplain = SIZE_MAX;
  }
  e->n_nary = r.n_fam[F_NARY];
  e->h_nary_ptr.resize(e->n_nary + 1);
  e->h_nary_kind.resize(e->n_nary);
  e->h_nary_ops.resize(r.n_nary_ops);
  e->h_tree_ptr.resize(e->n_nary);
  e->h_tree_nodes.resize(r.n_tree_ints);
  e->nary_uploaded = std::min(e->nary_uploaded, e->n_nary);
  e->nary_active_set = std::min(e->nary_active_set, e->n_nary);
  e->prop_ref.resize(r.n_props);
}

// pcp_engine_fork: the static arrays move into a refcounted block; this engine keeps views of them
// and puts its tail (everything at and behind n_static) into arrays of its own.
void share_static(pcp_engine* e) {
  if (e->shared) return;
  auto blk = std::make_shared<StaticBlock>();
  for (int f = 0; f < 3; ++f) {
    HostFamily& hf = e->fam[f];
    blk->desc[f] = hf.d_desc;  // (plain structs: pointer + capacity; the block owns them from here on)
    hf.tail_base = hf.n_static;
    hf.uploaded = std::min(hf.uploaded, hf.n_static);  // the tail is uploaded again, into d_tdesc
  }
  blk->descB = e->fam[F_TER].d_descB;
  blk->cdesc = e->fam[F_BIN].d_cdesc;
  blk->adj_ptr = e->d_adj_ptr;
  blk->adj = e->d_adj;
  e->shared = blk;
}
// Before anything that rewrites the static part (CSR rebuild, a restore below the shared prefix):
// let go of the shared block and upload everything again into arrays of this engine's own.
void unshare_static(pcp_engine* e) {
  if (!e->shared) return;
  e->shared.reset();
  for (int f = 0; f < 3; ++f) {
    HostFamily& hf = e->fam[f];
    hf.d_desc = DevBuf<int4>();
    hf.d_descB = DevBuf<int2>();
    hf.d_cdesc = DevBuf<uint2>();
    hf.uploaded = 0;
    hf.n_static = 0;
    hf.n_cdesc = 0;
    hf.tail_base = 0;
  }
  e->d_adj_ptr = DevBuf<int>();
  e->d_adj = DevBuf<uint32_t>();
  e->csr_built = false;
  e->at_fixpoint = false;
}

// Static reactor: CSR var -> refs of the propagators that depend on it, rows grouped by
// family (so the lanes of a warp expanding a row stay on one code path).
void build_csr(pcp_engine* e) {
  const size_t V = e->V;
  if (std::getenv("PCP_DEBUG_ALLOC")) std::fprintf(stderr, "[pcp csr] build V=%zu\n", V);
  std::vector<int> ptr(V + 1, 0);
  // dependencies of one lowered operand: the variable itself, or every term of a sum view
  auto dep = [&](int var, unsigned ref, auto&& fn) { for_each_dep_var(e, var, [&](int v) { fn(v, ref); }); };
  auto x_of = [](const int4& d) {
    unsigned v = (unsigned)d.x & kConstVar28;
    return v < kSumBase28 ? (int)v : (v == kConstVar28 ? -1 : -2 - (int)(v - kSumBase28));
  };
  auto for_each_var = [&](auto&& fn) {
    {
      const HostFamily& f = e->fam[F_BIN];
      for (size_t s = 0; s < f.n; ++s) {
        const int4& d = f.desc[s];
        dep(x_of(d), make_ref(F_BIN, (unsigned)s), fn);
        dep(d.z, make_ref(F_BIN, (unsigned)s), fn);
      }
    }
    {
      const HostFamily& f = e->fam[F_TER];
      for (size_t s = 0; s < f.n; ++s) {
        const int4& d = f.desc[s];
        dep(x_of(d), make_ref(F_TER, (unsigned)s), fn);
        dep(d.z, make_ref(F_TER, (unsigned)s), fn);
        dep(f.descB[s].x, make_ref(F_TER, (unsigned)s), fn);
      }
    }
    {
      const HostFamily& f = e->fam[F_DJ];
      std::vector<int> vars;
      for (size_t s = 0; s < f.n; ++s) {
        const int ops[6] = {f.desc[3 * s].x, f.desc[3 * s].z, f.desc[3 * s + 1].x,
                            f.desc[3 * s + 1].z, f.desc[3 * s + 2].x, f.desc[3 * s + 2].z};
        vars.clear();
        for (int o : ops) for_each_dep_var(e, o, [&](int v) { vars.push_back(v); });
        std::sort(vars.begin(), vars.end());  // sort + dedup union of the children (disjunction.rs:118-129)
        vars.erase(std::unique(vars.begin(), vars.end()), vars.end());
        for (int v : vars) fn(v, make_ref(F_DJ, (unsigned)s));
      }
    }
  };
  for_each_var([&](int v, unsigned) { ++ptr[v + 1]; });
  for (size_t v = 0; v < V; ++v) ptr[v + 1] += ptr[v];
  std::vector<uint32_t> adj((size_t)ptr[V]);
  std::vector<int> cur(ptr.begin(), ptr.end() - 1);
  for_each_var([&](int v, unsigned ref) { adj[(size_t)cur[v]++] = ref; });
  e->d_adj_ptr.reserve(V + 1, e->stream);
  e->d_adj.reserve(std::max<size_t>(adj.size(), 1), e->stream);
  CUDA_CHECK(cudaMemcpyAsync(e->d_adj_ptr.p, ptr.data(), (V + 1) * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  if (!adj.empty())
    CUDA_CHECK(cudaMemcpyAsync(e->d_adj.p, adj.data(), adj.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
  CUDA_CHECK(cudaStreamSynchronize(e->stream));  // host vectors go out of scope
  for (int f = 0; f < 3; ++f) { e->fam[f].n_static = e->fam[f].n; e->fam[f].static_kind_mask = e->fam[f].kind_mask; }
  // Compact stream for the lean binary sweep: when every static binary descriptor is an XNeqY
  // over plain variables with 16-bit variable ids and offsets (the n-queens / pairwise-distinct
  // stores), the sweep reads an 8-byte copy -- {xvar | yvar << 16, (u16)xoff | yoff << 16} --
  // and half the bytes cross HBM / L2.  The 16-byte descriptors stay the reference copy (rows,
  // tail, generic loop).
  {
    HostFamily& hb = e->fam[F_BIN];
    hb.n_cdesc = 0;
    const bool no_compact = std::getenv("PCP_NO_COMPACT") != nullptr;  // measurement / test switch (read at every reactor build)
    // (C5, 600 MB of descriptors: -20 % per node.  C2, 24 MB and L2-resident: a node alone on the GPU
    // loses 0.7 us to the unpacking, the contexts of a batched launch -- every SM sweeping at once --
    // gain 4-6 % from the halved L2 traffic; the batch is the configuration that counts.  Interval
    // domains only: the IntervalSet sweep has its witness probe in the 16-byte loop.)
    // word 0 = (xvar * 8) | (yvar * 8) << 16 (byte offsets into the domain array: V < 8192, which every
    // store with a shared-memory snapshot satisfies); word 1 = (u16) xoff | (yoff - xoff) << 16
    bool ok = !no_compact && !e->set_mode && hb.n > 0 && V < 8192 && hb.first_nonplain >= hb.n &&
              hb.static_kind_mask == (1 << B_NEQ);
    std::vector<uint2> cd;
    if (ok) {
      cd.resize(hb.n);
      for (size_t i = 0; i < hb.n && ok; ++i) {
        const int4& d = hb.desc[i];
        const long long delta = (long long)d.w - (long long)d.y;
        ok = d.y >= -32768 && d.y <= 32767 && delta >= -32768 && delta <= 32767;
        const unsigned xv = (unsigned)d.x & 0xffffu, yv = (unsigned)d.z & 0xffffu;
        cd[i] = make_uint2((xv * 8u) | ((yv * 8u) << 16), ((unsigned)d.y & 0xffffu) | ((unsigned)(int)delta << 16));
      }
    }
    if (ok) {
      hb.d_cdesc.reserve(hb.n, e->stream);
      CUDA_CHECK(cudaMemcpyAsync(hb.d_cdesc.p, cd.data(), hb.n * sizeof(uint2), cudaMemcpyHostToDevice, e->stream));
      CUDA_CHECK(cudaStreamSynchronize(e->stream));
      hb.n_cdesc = hb.n;
    }
  }
  e->csr_built = true;
}

size_t nary_smem_bytes(const pcp_engine* e) {
  if (e->n_nary == 0) return 0;
  size_t k = (size_t)e->nary_max_k;
  size_t tab = 4;
  while (tab < 2 * k) tab <<= 1;
  return k * 16 + tab * 4 + (e->set_mode ? k * 4 : 0);  // (set variants: + the list of newly assigned values)
}

// Everything except the few "inline" propagators is brought to the device here: new
// variables, bulk descriptor uploads, buffer growth, CSR (re)build.  Returns the launch
// parameters with the node prologue filled in; up to kMaxInline freshly posted tail
// propagators ride in the kernel parameters instead of an H2D copy.
Params prepare(pcp_engine* e) {
  const size_t V = e->V;
  // --- variables
  ensure_var_capacity(e, std::max<size_t>(V, 1));
  if (e->V_uploaded < V) {
    size_t n = V - e->V_uploaded;
    CUDA_CHECK(cudaMemcpyAsync(e->d_dom() + e->V_uploaded, e->h_dom_pending.data(), n * sizeof(int2),
                               cudaMemcpyHostToDevice, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    e->h_dom_pending.clear();
    size_t oldV = e->V_uploaded;
    if (e->set_mode) {
      int mn = INT32_MAX, mx = INT32_MIN;
      for (size_t v = oldV; v < V; ++v) { mn = std::min(mn, e->h_dom_init[v].x); mx = std::max(mx, e->h_dom_init[v].y); }
      if (e->set_W == 0) {  // the window is fixed by the first variables
        const long long words = ((long long)mx - mn + 1 + 31) / 32;
        if (words > 4096) PCP_FAIL(PCP_ERR_UNSUPPORTED, "IntervalSet domains spanning more than 131072 values are not supported");
        e->set_base = mn;
        e->set_W = (int)((words + 1) & ~1ll);  // even: a label slot stays 8-byte aligned
      } else if (mn < e->set_base || (long long)mx >= (long long)e->set_base + 32ll * e->set_W) {
        PCP_FAIL(PCP_ERR_UNSUPPORTED, "IntervalSet variable outside the value window fixed by the first variables");
      }
      // every value of the window starts in the set: bits outside the bounds are don't-care
      e->d_bits.reserve(V * (size_t)e->set_W, e->stream, oldV * (size_t)e->set_W);
      fill_u32(e, e->d_bits.p + oldV * (size_t)e->set_W, 0xffffffffu, (V - oldV) * (size_t)e->set_W);
    }
    e->V_uploaded = V;
    e->mirror_valid = false;
    ++e->dom_version;
    // the dirty sets: three bit sets over the variables, empty between launches
    (void)oldV;
    e->dirty_words = (V + 31) / 32;
    e->d_dirty_bits.reserve(3 * e->dirty_words, e->stream);
    CUDA_CHECK(cudaMemsetAsync(e->d_dirty_bits.p, 0, e->d_dirty_bits.cap * sizeof(uint32_t), e->stream));
    // the label stack is laid out with stride V: a label taken with fewer variables cannot be
    // restored after more were allocated
    if (e->stack_stride != e->slot_stride()) { PCP_REQUIRE(e->labels.empty(), "variables allocated while labels are live"); e->stack_stride = e->slot_stride(); }
    e->csr_built = false;  // adj_ptr has V+1 entries
    for (int f = 0; f < 3; ++f) e->fam[f].n_static = 0;
  }
  // --- reactor: (re)build when there are propagators outside the CSR beyond the tail limit
  size_t tail = 0, pending = 0;
  for (int f = 0; f < 3; ++f) { tail += e->fam[f].n - e->fam[f].n_static; pending += e->fam[f].n - e->fam[f].uploaded; }
  const bool rebuild = (!e->csr_built && tail > 0) || tail > e->tail_limit;
  if (rebuild && e->shared) {  // the static part is about to change: no longer the shared one
    unshare_static(e);
    pending = 0;
    for (int f = 0; f < 3; ++f) pending += e->fam[f].n;
  }
  // --- descriptors
  const bool use_inline = !rebuild && pending > 0 && pending <= (size_t)kMaxInline;
  size_t total = 0;
  for (int f = 0; f < 3; ++f) {
    HostFamily& hf = e->fam[f];
    const size_t w = (size_t)hf.width;
    if (e->shared) {
      // shared static model: the tail [tail_base, n) lives in this engine's own arrays
      const size_t t0 = hf.tail_base, have = hf.uploaded - t0, need = hf.n - t0;
      if (need * w > hf.d_tdesc.cap) hf.d_tdesc.reserve((need + e->tail_limit) * w, e->stream, have * w);
      if (f == F_TER && need > hf.d_tdescB.cap) hf.d_tdescB.reserve(need + e->tail_limit, e->stream, have);
      if (!use_inline && need > have) {
        CUDA_CHECK(cudaMemcpyAsync(hf.d_tdesc.p + have * w, hf.desc.data() + hf.uploaded * w, (need - have) * w * sizeof(int4),
                                   cudaMemcpyHostToDevice, e->stream));
        if (f == F_TER)
          CUDA_CHECK(cudaMemcpyAsync(hf.d_tdescB.p + have, hf.descB.data() + hf.uploaded, (need - have) * sizeof(int2),
                                     cudaMemcpyHostToDevice, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));  // pageable source
      }
    } else if (use_inline) {
      if (hf.desc.size() > hf.d_desc.cap) hf.d_desc.reserve(hf.desc.size() + e->tail_limit * w, e->stream, hf.uploaded * w);
      if (f == F_TER && hf.descB.size() > hf.d_descB.cap) hf.d_descB.reserve(hf.descB.size() + e->tail_limit, e->stream, hf.uploaded);
    } else {
      upload_range(e, hf.d_desc, hf.desc, hf.uploaded * w, hf.n * w);
      if (f == F_TER) upload_range(e, hf.d_descB, hf.descB, hf.uploaded, hf.n);
    }
    // (room for a whole tail: the branching constraints of a search are posted one by one, and a
    // reallocation -- cudaMalloc, copy, device synchronisation -- in the middle of it costs milliseconds)
    if ((hf.n + 31) / 32 + 1 > hf.d_active.cap)
      reserve_zeroed(e, hf.d_active, (hf.n + e->tail_limit + 31) / 32 + 2, (hf.active_set + 31) / 32);
    if (hf.n > hf.d_stamp.cap) {
      size_t valid = std::min(hf.d_stamp.cap, hf.active_set);
      hf.d_stamp.reserve(hf.n + e->tail_limit, e->stream, valid);
      fill_u32(e, hf.d_stamp.p + valid, 0u, hf.d_stamp.cap - valid);
    }
    if (!hf.d_desc.p && !e->shared) hf.d_desc.reserve(1, e->stream);
    if (!hf.d_descB.p && !e->shared) hf.d_descB.reserve(1, e->stream);
    if (e->shared && !hf.d_tdesc.p) hf.d_tdesc.reserve(1, e->stream);
    if (e->shared && !hf.d_tdescB.p) hf.d_tdescB.reserve(1, e->stream);
    if (!hf.d_stamp.p) { hf.d_stamp.reserve(1, e->stream); fill_u32(e, hf.d_stamp.p, 0u, hf.d_stamp.cap); }
    total += hf.n;
  }
  if (e->nary_uploaded < e->n_nary) {
    size_t ops_from = (size_t)e->h_nary_ptr[e->nary_uploaded];
    upload_range(e, e->d_nary_ops, e->h_nary_ops, ops_from, e->h_nary_ops.size());
    size_t pfrom = e->nary_uploaded == 0 ? 0 : e->nary_uploaded + 1;
    upload_range(e, e->d_nary_ptr, e->h_nary_ptr, pfrom, e->n_nary + 1);
    upload_range(e, e->d_nary_kind, e->h_nary_kind, e->nary_uploaded, e->n_nary);
    upload_range(e, e->d_tree_ptr, e->h_tree_ptr, e->nary_uploaded, e->n_nary);
    {  // the nodes of the trees among the new propagators
      size_t first = e->h_tree_nodes.size();
      for (size_t q = e->nary_uploaded; q < e->n_nary; ++q)
        if (e->h_tree_ptr[q] >= 0) { first = (size_t)e->h_tree_ptr[q]; break; }
      upload_range(e, e->d_tree_nodes, e->h_tree_nodes, first, e->h_tree_nodes.size());
    }
    e->nary_uploaded = e->n_nary;
  }
  reserve_zeroed(e, e->d_nary_active, (e->n_nary + 31) / 32 + 1, (e->nary_active_set + 31) / 32);
  total += e->n_nary;
  if (e->sums_uploaded < e->sums.size()) {  // Sum views are append-only (not part of a label)
    upload_range(e, e->d_sum_terms, e->h_sum_terms, (size_t)e->h_sum_ptr[e->sums_uploaded], e->h_sum_terms.size());
    upload_range(e, e->d_sum_ptr, e->h_sum_ptr, e->sums_uploaded == 0 ? 0 : e->sums_uploaded + 1, e->sums.size() + 1);
    e->sums_uploaded = e->sums.size();
  }
  if (total + 64 > e->d_trail.cap) e->d_trail.reserve(total + e->tail_limit + 64, e->stream, e->trail_len);

  Params P;
  std::memset(&P, 0, sizeof(P));
  // --- inline propagators (their descriptors are written by the kernel prologue)
  if (use_inline) {
    for (int f = 0; f < 3; ++f) {
      HostFamily& hf = e->fam[f];
      for (size_t s = hf.uploaded; s < hf.n; ++s) {
        InlineProp& ip = P.inl[P.n_inline++];
        ip.fam = (unsigned)f;
        ip.slot = (int)s;
        if (f == F_BIN) ip.q[0] = hf.desc[s];
        else if (f == F_TER) { ip.q[0] = hf.desc[s]; ip.q[1] = make_int4(hf.descB[s].x, hf.descB[s].y, 0, 0); }
        else { ip.q[0] = hf.desc[3 * s]; ip.q[1] = hf.desc[3 * s + 1]; ip.q[2] = hf.desc[3 * s + 2]; }
      }
    }
  }
  for (int f = 0; f < 3; ++f) e->fam[f].uploaded = e->fam[f].n;
  if (rebuild) {
    build_csr(e);
    e->at_fixpoint = false;
  }
  if (!e->csr_built) {  // no propagators yet: still need a valid adj_ptr
    e->d_adj_ptr.reserve(V + 1, e->stream);
    e->d_adj.reserve(1, e->stream);
    CUDA_CHECK(cudaMemsetAsync(e->d_adj_ptr.p, 0, (V + 1) * sizeof(int), e->stream));
  }

  // --- launch parameters
  P.result = e->d_result();
  P.dom = e->d_dom();
  P.V = (int)V;
  bool sync0 = false;
  for (int f = 0; f < 3; ++f) {
    Family& df = P.fam[f];
    HostFamily& hf = e->fam[f];
    df.desc = hf.d_desc.p;
    df.descB = hf.d_descB.p;
    df.tdesc = e->shared ? hf.d_tdesc.p - (ptrdiff_t)(hf.tail_base * (size_t)hf.width) : hf.d_desc.p;
    df.tdescB = e->shared ? hf.d_tdescB.p - (ptrdiff_t)hf.tail_base : hf.d_descB.p;
    df.active = hf.d_active.p;
    df.stamp = hf.d_stamp.p;
    df.n = (int)hf.n;
    df.n_static = (int)hf.n_static;
    df.all_plain = hf.first_nonplain >= hf.n_static ? 1 : 0;
    df.kind_mask = hf.static_kind_mask;
    df.cdesc = (f == F_BIN && hf.n_cdesc >= hf.n_static && hf.n_static > 0) ? hf.d_cdesc.p : nullptr;
    P.new_first[f] = (int)hf.active_set;
    P.new_last[f] = (int)hf.n;
    if (hf.active_set < hf.n && hf.active_set < hf.n_static) sync0 = true;  // bits other CTAs will read
    hf.active_set = hf.n;
  }
  P.nary_ptr = e->d_nary_ptr.p;
  P.nary_ops = e->d_nary_ops.p;
  P.nary_kind = e->d_nary_kind.p;
  P.tree_ptr = e->d_tree_ptr.p;
  P.tree_nodes = e->d_tree_nodes.p;
  P.nary_active = e->d_nary_active.p;
  P.nary_active_w = e->d_nary_active.p;
  P.n_nary = (int)e->n_nary;
  P.nary_max_k = e->nary_max_k;
  P.new_first[F_NARY] = (int)e->nary_active_set;
  P.new_last[F_NARY] = (int)e->n_nary;
  if (e->nary_active_set < e->n_nary) sync0 = true;
  e->nary_active_set = e->n_nary;
  P.sum_ptr = e->d_sum_ptr.p;
  P.sum_terms = e->d_sum_terms.p;
  P.adj_ptr = e->d_adj_ptr.p;
  P.adj = e->d_adj.p;
  P.dirty_bits = e->d_dirty_bits.p;
  P.dirty_words = (int)e->dirty_words;
  P.trail = e->d_trail.p;
  P.ctl = e->d_ctl;
  P.bits = e->d_bits.p;
  P.bits_W = e->set_W;
  P.bits_base = e->set_base;
  P.max_iterations = e->max_iterations;
  P.gen0 = e->bar_gen;
  P.epoch0 = e->epoch;
  if (e->pending_restore) {
    P.restore_from = e->d_stack.p + e->pending_restore_label * e->stack_stride;
    e->mirror_valid = false;
    ++e->dom_version;
    sync0 = true;
  }
  if (e->pending_trail_undo) {
    P.do_trail = 1;
    P.trail_keep = e->pending_trail_keep;
    if (e->trail_len > e->pending_trail_keep) sync0 = true;  // bits are actually re-set
    e->trail_len = std::min(e->trail_len, e->pending_trail_keep);
  }
  P.sync0 = sync0 ? 1 : 0;
  e->pending_restore = false;
  e->pending_trail_undo = false;
  return P;
}

bool prologue_pending(const pcp_engine* e) {
  if (e->V_uploaded < e->V || e->pending_restore || e->pending_trail_undo) return true;
  for (int f = 0; f < 3; ++f)
    if (e->fam[f].uploaded < e->fam[f].n || e->fam[f].active_set < e->fam[f].n) return true;
  return e->nary_uploaded < e->n_nary || e->nary_active_set < e->n_nary;
}

__global__ void __launch_bounds__(1024, 1) pcp_prologue_kernel(const __grid_constant__ Params P) {
  full::node_prologue(P, threadIdx.x, blockDim.x);  // single CTA
  __syncthreads();
  if (threadIdx.x == 0) full::node_prologue_finish(P);
}
__global__ void __launch_bounds__(1024, 1) pcp_prologue_set_kernel(const __grid_constant__ Params P) {
  full_set::node_prologue(P, threadIdx.x, blockDim.x);  // (restores the bit sets as well)
  __syncthreads();
  if (threadIdx.x == 0) full_set::node_prologue_finish(P);
}
// Cardinality::size() of every IntervalSet domain (when the fixpoint kernel did not leave them
// in the host mirror: stores too large for the zero-copy result, reads after a restore).
__global__ void pcp_set_sizes_kernel(const int2* dom, const uint32_t* bits, int W, int base, int V, uint32_t* out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int2 d = dom[v];
  full_set::SetW sw{const_cast<uint32_t*>(bits), W, base};
  out[v] = d.x > d.y ? 0u : full_set::sb_count(sw, v, d.x, d.y);
}

// Bring the device state up to date without running a fixpoint (pcp_domains_read /
// pcp_label / pcp_active_read right after a restore or an alloc).
void sync_device_state(pcp_engine* e) {
  if (!prologue_pending(e)) return;
  Params P = prepare(e);
  if (e->set_mode) pcp_prologue_set_kernel<<<1, 1024, 0, e->stream>>>(P);
  else pcp_prologue_kernel<<<1, 1024, 0, e->stream>>>(P);
  CUDA_CHECK(cudaGetLastError());
}

// Kernel variants: `binonly` for stores whose propagators are all binary (no ternary,
// disjunction or n-ary code in the kernel -- see pcp_body.cuh), `full` otherwise.
bool bin_only_store(const pcp_engine* e) { return e->fam[F_TER].n == 0 && e->fam[F_DJ].n == 0 && e->n_nary == 0; }
const void* fixpoint_set_fn(bool bin_only, bool smem_dom) {
  if (bin_only) return smem_dom ? (const void*)binonly_set::pcp_fixpoint_kernel<true> : (const void*)binonly_set::pcp_fixpoint_kernel<false>;
  return smem_dom ? (const void*)full_set::pcp_fixpoint_kernel<true> : (const void*)full_set::pcp_fixpoint_kernel<false>;
}
const void* fixpoint_fn(bool bin_only, bool smem_dom) {
  if (bin_only) return smem_dom ? (const void*)binonly::pcp_fixpoint_kernel<true> : (const void*)binonly::pcp_fixpoint_kernel<false>;
  return smem_dom ? (const void*)full::pcp_fixpoint_kernel<true> : (const void*)full::pcp_fixpoint_kernel<false>;
}
const void* fixpoint_batch_fn(bool set_mode, bool bin_only) {
  if (set_mode) return bin_only ? (const void*)binonly_set::pcp_fixpoint_batch_kernel : (const void*)full_set::pcp_fixpoint_batch_kernel;
  return bin_only ? (const void*)binonly::pcp_fixpoint_batch_kernel : (const void*)full::pcp_fixpoint_batch_kernel;
}
const void* burst_set_fn(bool bin_only, bool smem_dom) {
  if (bin_only) return smem_dom ? (const void*)binonly_set::pcp_burst_kernel<true> : (const void*)binonly_set::pcp_burst_kernel<false>;
  return smem_dom ? (const void*)full_set::pcp_burst_kernel<true> : (const void*)full_set::pcp_burst_kernel<false>;
}
const void* burst_fn(bool bin_only, bool smem_dom) {
  if (bin_only) return smem_dom ? (const void*)binonly::pcp_burst_kernel<true> : (const void*)binonly::pcp_burst_kernel<false>;
  return smem_dom ? (const void*)full::pcp_burst_kernel<true> : (const void*)full::pcp_burst_kernel<false>;
}

// The persistent kernels need every CTA co-resident (they meet at a device-wide barrier):
// grid <= number of SMs at one CTA per SM.  A cooperative launch makes the driver check
// that; PCP_LAUNCH=plain uses an ordinary launch (same grid; valid while the engine has the
// device to itself) -- kept as a measurement switch.
void launch_persistent(const void* fn, int grid, void** args, size_t smem, cudaStream_t stream) {
  static const bool plain = [] { const char* v = std::getenv("PCP_LAUNCH"); return v && std::strcmp(v, "plain") == 0; }();
  if (plain) CUDA_CHECK(cudaLaunchKernel(fn, dim3(grid), dim3(kThreads), args, smem, stream));
  else CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args, smem, stream));
}

// CTAs a persistent launch of this engine may occupy: all SMs, or the share it was given when
// several engines run their fixpoints side by side on one GPU (pcp_set_grid_limit)
int launch_ctas(const pcp_engine* e) {
  static const int env = [] { const char* v = std::getenv("PCP_MAX_CTAS"); return v ? std::atoi(v) : 0; }();  // measurement switch
  int n = e->num_sms;
  if (e->grid_limit > 0) n = std::min(n, e->grid_limit);
  if (env > 0) n = std::min(n, env);
  return std::max(n, 1);
}

struct HostProf {
  double prepare = 0, launch = 0, wait = 0, total = 0;
  unsigned long long n = 0;
  ~HostProf() {
    if (n && std::getenv("PCP_HOSTPROF"))
      std::fprintf(stderr, "[pcp hostprof] %llu fixpoint calls: prepare %.2f us, launch API %.2f us, wait %.2f us, total %.2f us per call\n",
                   n, 1e6 * prepare / n, 1e6 * launch / n, 1e6 * wait / n, 1e6 * total / n);
  }
};
static HostProf g_hostprof;
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// One fixpoint = fixpoint_prepare (uploads, node prologue parameters, launch geometry) +
// fixpoint_fire (the launch) + fixpoint_wait (result header and domains back, host bookkeeping).  pcp_consistency runs them back
// to back; pcp_consistency_batch launches several engines' fixpoints before it waits for any, so
// that their persistent grids run side by side on the GPU.
void fixpoint_prepare(pcp_engine* e) {
  PCP_REQUIRE(!e->inflight.active, "a fixpoint of this engine is already in flight");
  const double hp0 = now_s();
  Params P = prepare(e);
  const double hp1 = now_s();
  const size_t V = e->V;
  const bool incremental = (e->flags & PCP_FLAG_INCREMENTAL) && e->at_fixpoint;
  P.full_sweep = incremental ? 0 : 1;

  // epoch wrap: stamps are compared for equality with epochs of this launch only
  if (e->epoch > 0x7f000000u) {
    for (int f = 0; f < 3; ++f) fill_u32(e, e->fam[f].d_stamp.p, 0u, e->fam[f].d_stamp.cap);
    e->epoch = 1;
    P.epoch0 = 1;
  }
  if (incremental && !e->host_dirty.empty()) {
    // seed the worklist of iteration 0 with the variables narrowed by the host
    std::sort(e->host_dirty.begin(), e->host_dirty.end());
    e->host_dirty.erase(std::unique(e->host_dirty.begin(), e->host_dirty.end()), e->host_dirty.end());
    size_t n = e->host_dirty.size();
    int* st = static_cast<int*>(stage(e, n * sizeof(int)));
    std::memcpy(st, e->host_dirty.data(), n * sizeof(int));
    e->d_seed_list.reserve(n, e->stream);
    CUDA_CHECK(cudaMemcpyAsync(e->d_seed_list.p, st, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    pcp_seed_dirty_kernel<<<(int)((n + 255) / 256), 256, 0, e->stream>>>(e->d_seed_list.p, (int)n, e->d_dirty_bits.p);
    CUDA_CHECK(cudaGetLastError());
    P.seed_dirty = (int)n;
  }
  e->host_dirty.clear();

  // label slot for the epilogue snapshot (Branch::distribute labels right after an Unknown node)
  e->snapshot_valid = false;
  const bool want_snapshot = V > 0 && V <= 16384 && e->labels.size() < e->max_labels && e->stack_stride == e->slot_stride();
  if (want_snapshot) {
    size_t idx = e->labels.size();
    if ((idx + 1) * e->stack_stride > e->d_stack.cap) {
      const int2* old = e->d_stack.p;
      // grow in big steps: a reallocation costs a cudaMalloc/cudaFree pair on the search path
      size_t want = std::max<size_t>((idx + 1) * 2, std::min<size_t>(e->max_labels, 256));
      e->d_stack.reserve(want * e->stack_stride, e->stream, idx * e->stack_stride);
      if (P.restore_from) P.restore_from = e->d_stack.p + (P.restore_from - old);
    }
    P.snapshot_to = e->d_stack.p + idx * e->stack_stride;
  }

  // launch geometry: one CTA per SM, fewer for small stores (cheaper barrier)
  size_t total = e->n_nary * 4096;
  for (int f = 0; f < 3; ++f) total += e->fam[f].n;
  int grid = (int)std::min<size_t>((size_t)launch_ctas(e), std::max<size_t>(1, (total + 4095) / 4096));
  size_t nary_bytes = nary_smem_bytes(e);
  PCP_REQUIRE(nary_bytes <= (size_t)kRingBytes, "Distinct too wide for shared memory");
  size_t dom_bytes = (V * 8 + 15) & ~size_t(15);
  bool smem_dom = V > 0 && (size_t)kRingBytes + dom_bytes + e->static_smem + 256 <= (size_t)e->max_smem_optin;
  size_t smem = (size_t)kRingBytes + (smem_dom ? dom_bytes : 0);
  P.smem_dom = smem_dom ? 1 : 0;
  {  // per-CTA bitmap of the variables a sweep narrows (flushed into the dirty set once per sweep):
     // for the stores too large for the snapshot, whose first sweeps narrow most variables many
     // times over (C4: -12 %); a store with a snapshot narrows little per sweep, and there the
     // flush costs more than the scattered reductions it saves (C2: +2 %)
    const size_t bm_bytes = (e->dirty_words * 4 + 15) & ~size_t(15);
    P.dirty_bm_off = 0;
    if (V > 0 && !smem_dom && smem + bm_bytes + e->static_smem + 256 <= (size_t)e->max_smem_optin) {
      P.dirty_bm_off = (int)smem;
      smem += bm_bytes;
    }
  }
  const void* fn = e->set_mode ? fixpoint_set_fn(bin_only_store(e), smem_dom) : fixpoint_fn(bin_only_store(e), smem_dom);

  static const bool trace_on = std::getenv("PCP_TRACE") != nullptr;
  if (trace_on) {
    if (!e->d_trace) CUDA_CHECK(cudaMalloc(&e->d_trace, kTraceWords * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMemsetAsync(e->d_trace, 0, kTraceWords * sizeof(unsigned long long), e->stream));
    P.trace = e->d_trace;
  }
  // The result header (+ the domains when small) comes back without a copy or a stream
  // synchronisation: the kernel's epilogue stores them into the mapped pinned mirror and
  // publishes a sequence number last; the host polls that word.  (Timed or traced launches
  // synchronise anyway.)
  const bool eager_dom = V * sizeof(int2) <= (64u << 10);
  const bool zero_copy = !trace_on && !e->timing && e->h_block_dev != nullptr;
  if (zero_copy) {
    P.host_result = reinterpret_cast<Result*>(e->h_block_dev);
    P.host_dom = eager_dom ? 1 : 0;
    P.host_seq = ++e->host_seq ? e->host_seq : ++e->host_seq;  // never 0
    if (e->set_mode && eager_dom) P.host_sizes = e->h_sizes_dev();
  }
  e->sizes_valid = false;
  pcp_engine::Inflight& f = e->inflight;
  f.prepared = true;
  f.fused = false;
  f.P = P;
  f.fn = fn;
  f.smem = smem;
  f.grid = grid;
  f.want_snapshot = want_snapshot;
  f.eager_dom = eager_dom;
  f.zero_copy = zero_copy;
  f.hp0 = hp0; f.hp1 = hp1;
}

// the launch itself: nothing but the (cooperative) launch call between the two timing events
void fixpoint_fire(pcp_engine* e) {
  pcp_engine::Inflight& f = e->inflight;
  PCP_REQUIRE(f.prepared && !f.active, "fixpoint not prepared");
  f.hp2 = now_s();
  if (e->timing) CUDA_CHECK(cudaEventRecord(e->ev0, e->stream));
  void* args[] = {&f.P};
  launch_persistent(f.fn, f.grid, args, f.smem, e->stream);
  if (e->timing) CUDA_CHECK(cudaEventRecord(e->ev1, e->stream));
  f.hp3 = now_s();
  ++e->dom_version;
  f.prepared = false;
  f.active = true;
}

void fixpoint_launch(pcp_engine* e) {
  fixpoint_prepare(e);
  fixpoint_fire(e);
}

void fixpoint_wait(pcp_engine* e, int32_t* status, pcp_stats* stats) {
  pcp_engine::Inflight& f = e->inflight;
  PCP_REQUIRE(f.active, "no fixpoint of this engine is in flight");
  f.active = false;
  const Params& P = f.P;
  const size_t V = e->V;
  const int grid = f.grid;
  const bool want_snapshot = f.want_snapshot, eager_dom = f.eager_dom, zero_copy = f.zero_copy;
  const double hp0 = f.hp0, hp1 = f.hp1, hp2 = f.hp2, hp3 = f.hp3;
  static const bool trace_on = std::getenv("PCP_TRACE") != nullptr;
  if (zero_copy) {
    volatile unsigned* seq = &e->h_result()->seq;
    for (unsigned long long spins = 1; *seq != P.host_seq; ++spins) {
      if ((spins & 0xfffffull) == 0) {  // every ~1M polls: has the launch died?
        cudaError_t q = cudaStreamQuery(e->stream);
        if (q != cudaSuccess && q != cudaErrorNotReady) CUDA_CHECK(q);
        if (q == cudaSuccess && *seq != P.host_seq) PCP_FAIL(PCP_ERR_CUDA, "fixpoint kernel finished without publishing its result");
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  } else {
    // status header (+ domains when small) in one D2H copy
    size_t bytes = sizeof(Result) + (eager_dom ? V * sizeof(int2) : 0);
    CUDA_CHECK(cudaMemcpyAsync(e->h_block, e->d_block, bytes, cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
  }
  e->mirror_valid = eager_dom;
  e->sizes_valid = P.host_sizes != nullptr && !e->h_result()->failed;
  {
    const double hp4 = now_s();
    static unsigned long long skip = 0;
    if (++skip > 10) {  // (the first calls build the CSR and upload the store)
      g_hostprof.prepare += hp1 - hp0; g_hostprof.launch += hp3 - hp2; g_hostprof.wait += hp4 - hp3; g_hostprof.total += hp4 - hp0;
      ++g_hostprof.n;
    }
  }

  if (trace_on) {
    std::vector<unsigned long long> t(kTraceWords);
    CUDA_CHECK(cudaMemcpy(t.data(), e->d_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    {
      unsigned long long tb = ~0ull;
      for (int b = 0; b < grid; ++b) tb = std::min(tb, t[b * 8]);
      for (unsigned it = 0; it < e->h_result()->iterations && it < 32; ++it)
        std::fprintf(stderr, "[pcp trace]   iter %2u: barrier left at %8llu ns, dirty_in=%llu, decision=%llu%s\n", it,
                     t[8 * 256 + it * 4] - tb, t[8 * 256 + it * 4 + 1], t[8 * 256 + it * 4 + 2] & 7,
                     (t[8 * 256 + it * 4 + 2] & 8) ? " (sweep)" : "");
    }
    unsigned long long t0 = ~0ull;
    for (int b = 0; b < grid; ++b) t0 = std::min(t0, t[b * 8]);
    std::fprintf(stderr, "[pcp trace] grid=%d sync0=%d iters=%u (first CTA start at globaltimer %llu ns); per phase: min/avg/max ns since first CTA start\n",
                 grid, P.sync0, e->h_result()->iterations, t0);
    const char* names[7] = {"start", "after_prologue_sync", "snapshot+inline_done", "sweep_done", "barrier_arrive", "barrier_leave", "exit"};
    for (int k = 0; k < 7; ++k) {
      unsigned long long mn = ~0ull, mx = 0, sum = 0;
      for (int b = 0; b < grid; ++b) { unsigned long long v = t[b * 8 + k] - t0; mn = std::min(mn, v); mx = std::max(mx, v); sum += v; }
      std::fprintf(stderr, "[pcp trace]   %-22s %8llu %8llu %8llu\n", names[k], mn, sum / grid, mx);
    }
    {  // the stragglers of iteration 0: CTAs by time of their last sweep chunk
      std::vector<std::pair<unsigned long long, int>> by;
      for (int b = 0; b < grid; ++b) by.push_back({t[b * 8 + 3] - t0, b});
      std::sort(by.begin(), by.end());
      std::fprintf(stderr, "[pcp trace]   sweep_done by CTA: fastest");
      for (int i = 0; i < 4 && i < grid; ++i) std::fprintf(stderr, " %d:%llu", by[i].second, by[i].first);
      std::fprintf(stderr, "  median %llu  slowest", by[grid / 2].first);
      for (int i = std::max(0, grid - 8); i < grid; ++i) std::fprintf(stderr, " %d:%llu(start %llu)", by[i].second, by[i].first, t[by[i].second * 8] - t0);
      std::fprintf(stderr, "\n");
    }
    if (e->h_result()->iterations > 1 && t[kTraceSeq] > 0) {
      std::fprintf(stderr, "[pcp trace]   cta0 row-local sequence (id:ns since first CTA start):");
      for (unsigned long long i = 0; i < t[kTraceSeq] && i < 64; ++i)
        std::fprintf(stderr, " %llu:%llu", t[kTraceSeq + 2 + 2 * i], t[kTraceSeq + 3 + 2 * i] - t0);
      std::fprintf(stderr, "\n");
    }
    if (e->h_result()->iterations > 1) {
      const char* n1[6] = {"it1 start", "it1 compacted", "it1 refreshed", "it1 rows/sweep done", "it1 barrier_arrive", "it1 barrier_leave"};
      for (int k = 0; k < 6; ++k) {
        unsigned long long mn = ~0ull, mx = 0, sum = 0;
        int cnt = 0;
        for (int b = 0; b < grid; ++b) {
          unsigned long long raw = t[kTraceIter1 + b * 8 + k];
          if (!raw) continue;
          unsigned long long v = raw - t0; mn = std::min(mn, v); mx = std::max(mx, v); sum += v; ++cnt;
        }
        if (cnt) std::fprintf(stderr, "[pcp trace]   %-22s %8llu %8llu %8llu  (%d CTAs)  cta0 %llu\n", n1[k], mn, sum / cnt, mx, cnt,
                              t[kTraceIter1 + k] ? t[kTraceIter1 + k] - t0 : 0ull);
      }
    }
  }
  const Result& r = *e->h_result();
  if (r.decision == D_ITER_CAP) PCP_FAIL(PCP_ERR_CUDA, "fixpoint iteration cap reached");
  e->trail_len = r.trail_cnt;
  e->epoch = r.epoch;
  e->bar_gen = r.gen;
  unsigned long long props = r.propagations - e->props_total;
  e->props_total = r.propagations;
  // propagation/store.rs:250-256: False | True (every propagator entailed) | Unknown
  int st = r.failed ? PCP_FALSE : (r.trail_cnt == e->num_props() ? PCP_TRUE : PCP_UNKNOWN);
  e->at_fixpoint = !r.failed;
  if (want_snapshot && !r.failed) { e->snapshot_valid = true; e->snapshot_version = e->dom_version; }
  *status = st;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->propagations = props;
    stats->iterations = r.iterations;
    stats->active_props = (uint32_t)(e->num_props() - r.trail_cnt);
    stats->launches = f.fused ? 0u : 1u;
    if (e->timing && !f.fused) {
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
      stats->kernel_ms = ms;
    }
  }
}

// pcp_consistency_batch: can these prepared fixpoints share one launch?  Same kernel variant, same
// geometry, snapshot mode, one device, and room for every group at one CTA per SM.
bool batch_fusable(pcp_engine* const* engines, int n) {
  const char* sw = std::getenv("PCP_BATCH");  // measurement / test switch
  const bool off = sw && std::strcmp(sw, "unfused") == 0;
  static const bool trace_on = std::getenv("PCP_TRACE") != nullptr;
  if (off || trace_on || n < 2) return false;
  const pcp_engine* lead = engines[0];
  const pcp_engine::Inflight& f0 = lead->inflight;
  if (!f0.P.smem_dom || (long long)n * f0.grid > lead->num_sms) return false;
  for (int i = 0; i < n; ++i) {
    const pcp_engine* e = engines[i];
    const pcp_engine::Inflight& f = e->inflight;
    if (e->device != lead->device || f.fn != f0.fn || f.smem != f0.smem || f.grid != f0.grid || !f.prepared) return false;
    for (int j = 0; j < i; ++j)
      if (engines[j] == e) return false;
  }
  return true;
}

// the one launch for all of them, on the lead engine's stream; the other engines' streams are ordered
// before it (their uploads) and after it (their result copies)
void fire_fused(pcp_engine* const* engines, int n, bool timed) {
  pcp_engine* lead = engines[0];
  if (lead->batch_cap < n) {
    if (lead->d_batch) cudaFree(lead->d_batch);
    if (lead->h_batch) cudaFreeHost(lead->h_batch);
    lead->d_batch = lead->h_batch = nullptr;
    lead->batch_cap = 0;
    const int cap = std::max(n, 32);
    CUDA_CHECK(cudaMalloc(&lead->d_batch, (size_t)cap * sizeof(Params)));
    CUDA_CHECK(cudaHostAlloc(&lead->h_batch, (size_t)cap * sizeof(Params), cudaHostAllocDefault));
    lead->batch_cap = cap;
  }
  const double t0 = now_s();
  for (int i = 0; i < n; ++i) lead->h_batch[i] = engines[i]->inflight.P;
  for (int i = 1; i < n; ++i) {
    CUDA_CHECK(cudaEventRecord(engines[i]->ev_done, engines[i]->stream));
    CUDA_CHECK(cudaStreamWaitEvent(lead->stream, engines[i]->ev_done, 0));
  }
  CUDA_CHECK(cudaMemcpyAsync(lead->d_batch, lead->h_batch, (size_t)n * sizeof(Params), cudaMemcpyHostToDevice, lead->stream));
  const pcp_engine::Inflight& f0 = lead->inflight;
  const void* fn = fixpoint_batch_fn(lead->set_mode, bin_only_store(lead));
  const Params* batch = lead->d_batch;
  int group = f0.grid;
  void* args[] = {(void*)&batch, (void*)&group};
  if (timed) CUDA_CHECK(cudaEventRecord(lead->ev_batch0, lead->stream));
  launch_persistent(fn, n * group, args, f0.smem, lead->stream);
  if (timed) CUDA_CHECK(cudaEventRecord(lead->ev_batch1, lead->stream));
  CUDA_CHECK(cudaEventRecord(lead->ev_done, lead->stream));
  const double t1 = now_s();
  for (int i = 0; i < n; ++i) {
    pcp_engine* e = engines[i];
    if (i > 0) CUDA_CHECK(cudaStreamWaitEvent(e->stream, lead->ev_done, 0));
    pcp_engine::Inflight& f = e->inflight;
    f.hp2 = t0; f.hp3 = t1;
    ++e->dom_version;
    f.prepared = false;
    f.active = true;
    f.fused = true;
  }
}

void run_fixpoint(pcp_engine* e, int32_t* status, pcp_stats* stats) {
  fixpoint_launch(e);
  fixpoint_wait(e, status, stats);
}

void fetch_domains(pcp_engine* e) {
  sync_device_state(e);
  if (e->mirror_valid) return;
  e->sizes_valid = false;
  if (e->V) {
    CUDA_CHECK(cudaMemcpyAsync(e->h_dom(), e->d_dom(), e->V * sizeof(int2), cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
  }
  e->mirror_valid = true;
}

template <class F>
int guarded(pcp_engine* e, F&& f) {
  try {
    f();
    return PCP_OK;
  } catch (const Error& er) {
    if (e) e->err = er.msg;
    return er.code;
  } catch (const std::bad_alloc&) {
    if (e) e->err = "host allocation failed";
    return PCP_ERR_NOMEM;
  }
}

}  // namespace

extern "C" {

int pcp_engine_create(const pcp_config* cfg, pcp_engine** out) {
  if (!out) return PCP_ERR_INVALID;
  *out = nullptr;
  pcp_engine* e = new pcp_engine();
  int rc = guarded(e, [&] {
    e->device = cfg ? cfg->device : 0;
    e->flags = cfg ? cfg->flags : 0;
    e->set_mode = (e->flags & PCP_FLAG_INTERVAL_SET) != 0;
    if (cfg && cfg->max_labels) e->max_labels = cfg->max_labels;
    if (cfg && cfg->tail_limit) e->tail_limit = cfg->tail_limit;
    // engines running side by side need their streams on distinct hardware queues (default: 8);
    // only effective when set before the process creates its CUDA context, never overrides the user
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) PCP_FAIL(PCP_ERR_CUDA, "no CUDA device: the propagation engine has no CPU fallback");
    PCP_REQUIRE(e->device >= 0 && e->device < count, "device ordinal out of range");
    CUDA_CHECK(cudaSetDevice(e->device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major < 10) PCP_FAIL(PCP_ERR_CUDA, "device is not sm_100 (Blackwell); the kernels are built for sm_100a only");
    int coop = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device));
    if (!coop) PCP_FAIL(PCP_ERR_CUDA, "device does not support cooperative launch");
    e->num_sms = prop.multiProcessorCount;
    e->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    {
      // dynamic shared memory every persistent kernel may use = the opt-in maximum minus its own
      // static part; configured once (an engine never lowers what another engine relies on)
      const void* fns[24] = {(const void*)binonly::pcp_burst_batch_kernel, (const void*)full::pcp_burst_batch_kernel,
                             (const void*)binonly_set::pcp_burst_batch_kernel, (const void*)full_set::pcp_burst_batch_kernel,
                             fixpoint_batch_fn(false, false), fixpoint_batch_fn(false, true), fixpoint_batch_fn(true, false), fixpoint_batch_fn(true, true),
                             burst_set_fn(false, false), burst_set_fn(false, true), burst_set_fn(true, false), burst_set_fn(true, true),fixpoint_fn(false, false), fixpoint_fn(false, true), burst_fn(false, false), burst_fn(false, true),
                             fixpoint_fn(true, false),  fixpoint_fn(true, true),  burst_fn(true, false),  burst_fn(true, true),
                             fixpoint_set_fn(false, false), fixpoint_set_fn(false, true), fixpoint_set_fn(true, false),
                             fixpoint_set_fn(true, true)};
      size_t max_static = 0;
      for (const void* fn : fns) {
        cudaFuncAttributes fa;
        CUDA_CHECK(cudaFuncGetAttributes(&fa, fn));
        max_static = std::max(max_static, fa.sharedSizeBytes);
        CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)e->max_smem_optin - fa.sharedSizeBytes)));
      }
      e->static_smem = max_static;
    }
    CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreate(&e->ev0));
    CUDA_CHECK(cudaEventCreate(&e->ev1));
    CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_sleep, cudaEventDisableTiming | cudaEventBlockingSync));
    CUDA_CHECK(cudaEventCreate(&e->ev_batch0));
    CUDA_CHECK(cudaEventCreate(&e->ev_batch1));
    CUDA_CHECK(cudaMalloc(&e->d_ctl, sizeof(Control)));
    Control c;
    std::memset(&c, 0, sizeof(c));
    c.epoch = 1;
    CUDA_CHECK(cudaMemcpy(e->d_ctl, &c, sizeof(c), cudaMemcpyHostToDevice));
    e->fam[F_BIN].width = 1;
    e->fam[F_TER].width = 1;
    e->fam[F_DJ].width = 3;
    stage(e, 1 << 16);
  });
  if (rc != PCP_OK) {
    std::fprintf(stderr, "pcp_engine_create: %s\n", e->err.c_str());
    delete e;
    return rc;
  }
  *out = e;
  return PCP_OK;
}

int pcp_engine_fork(pcp_engine* parent, pcp_engine** out) {
  if (!parent || !out) return PCP_ERR_INVALID;
  *out = nullptr;
  // bring the parent's device state up to date, then share its static model
  int rc = guarded(parent, [&] {
    PCP_REQUIRE_NO_BURST(parent);
    PCP_REQUIRE(!parent->inflight.active, "a fixpoint of this engine is in flight");
    CUDA_CHECK(cudaSetDevice(parent->device));
    sync_device_state(parent);
    if (!parent->csr_built) {  // the reactor is built by the first fixpoint: run the prologue path that builds it
      size_t tail = 0;
      for (int f = 0; f < 3; ++f) tail += parent->fam[f].n - parent->fam[f].n_static;
      PCP_REQUIRE(tail == 0, "fork before the first pcp_consistency of a store with propagators");
    }
    CUDA_CHECK(cudaStreamSynchronize(parent->stream));
    share_static(parent);
  });
  if (rc != PCP_OK) return rc;
  pcp_engine* e = new pcp_engine();
  rc = guarded(e, [&] {
    const pcp_engine* p = parent;
    e->device = p->device;
    e->flags = p->flags;
    e->set_mode = p->set_mode;
    e->set_base = p->set_base;
    e->set_W = p->set_W;
    e->max_labels = p->max_labels;
    e->tail_limit = p->tail_limit;
    e->num_sms = p->num_sms;
    e->max_smem_optin = p->max_smem_optin;
    e->static_smem = p->static_smem;
    e->max_iterations = p->max_iterations;
    CUDA_CHECK(cudaSetDevice(e->device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreate(&e->ev0));
    CUDA_CHECK(cudaEventCreate(&e->ev1));
    CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_sleep, cudaEventDisableTiming | cudaEventBlockingSync));
    CUDA_CHECK(cudaEventCreate(&e->ev_batch0));
    CUDA_CHECK(cudaEventCreate(&e->ev_batch1));
    stage(e, 1 << 16);
    // ---- variables: the parent's current domains (and bit sets)
    e->V = e->V_uploaded = p->V;
    e->h_dom_init = p->h_dom_init;
    ensure_var_capacity(e, std::max<size_t>(e->V, 1));
    if (e->V) CUDA_CHECK(cudaMemcpyAsync(e->d_dom(), p->d_dom(), e->V * sizeof(int2), cudaMemcpyDeviceToDevice, e->stream));
    if (e->set_mode && e->V) {
      e->d_bits.reserve(e->V * (size_t)e->set_W, e->stream);
      CUDA_CHECK(cudaMemcpyAsync(e->d_bits.p, p->d_bits.p, e->V * (size_t)e->set_W * 4, cudaMemcpyDeviceToDevice, e->stream));
    }
    e->dirty_words = p->dirty_words;
    e->d_dirty_bits.reserve(std::max<size_t>(3 * e->dirty_words, 1), e->stream);
    CUDA_CHECK(cudaMemsetAsync(e->d_dirty_bits.p, 0, e->d_dirty_bits.cap * sizeof(uint32_t), e->stream));
    e->stack_stride = e->slot_stride();
    // ---- propagators: host descriptors copied, static device arrays shared, tail / active / stamps own
    e->shared = p->shared;
    for (int f = 0; f < 3; ++f) {
      const HostFamily& pf = p->fam[f];
      HostFamily& hf = e->fam[f];
      hf.desc = pf.desc;
      hf.descB = pf.descB;
      hf.n = pf.n;
      hf.n_static = pf.n_static;
      hf.tail_base = pf.tail_base;
      hf.uploaded = pf.tail_base;          // this engine's tail is uploaded by its first prepare()
      hf.active_set = pf.active_set;
      hf.first_nonplain = pf.first_nonplain;
      hf.kind_mask = pf.kind_mask;
      hf.static_kind_mask = pf.static_kind_mask;
      hf.width = pf.width;
      hf.n_cdesc = pf.n_cdesc;
      hf.d_desc = pf.d_desc;               // views of the shared block
      hf.d_descB = pf.d_descB;
      hf.d_cdesc = pf.d_cdesc;
      const size_t words = (hf.n + e->tail_limit + 31) / 32 + 2;
      hf.d_active.reserve(words, e->stream);
      fill_u32(e, hf.d_active.p, 0u, hf.d_active.cap);
      const size_t pw = std::min(words, (pf.active_set + 31) / 32);
      if (pw) CUDA_CHECK(cudaMemcpyAsync(hf.d_active.p, pf.d_active.p, pw * 4, cudaMemcpyDeviceToDevice, e->stream));
      hf.d_stamp.reserve(hf.n + e->tail_limit + 1, e->stream);
      fill_u32(e, hf.d_stamp.p, 0u, hf.d_stamp.cap);
    }
    e->d_adj_ptr = p->d_adj_ptr;
    e->d_adj = p->d_adj;
    e->csr_built = p->csr_built;
    e->h_nary_ptr = p->h_nary_ptr;
    e->h_nary_ops = p->h_nary_ops;
    e->h_nary_kind = p->h_nary_kind;
    e->h_tree_ptr = p->h_tree_ptr;
    e->h_tree_nodes = p->h_tree_nodes;
    e->n_nary = p->n_nary;
    e->nary_max_k = p->nary_max_k;
    e->nary_uploaded = 0;                  // small arrays: uploaded again by the first prepare()
    e->nary_active_set = p->nary_active_set;
    e->d_nary_active.reserve((e->n_nary + 31) / 32 + 1, e->stream);
    fill_u32(e, e->d_nary_active.p, 0u, e->d_nary_active.cap);
    if (p->nary_active_set)
      CUDA_CHECK(cudaMemcpyAsync(e->d_nary_active.p, p->d_nary_active.p, ((p->nary_active_set + 31) / 32) * 4, cudaMemcpyDeviceToDevice, e->stream));
    e->prop_ref = p->prop_ref;
    e->sums = p->sums;
    e->h_sum_ptr = p->h_sum_ptr;
    e->h_sum_terms = p->h_sum_terms;
    e->sums_uploaded = 0;
    // ---- trail and control block
    e->trail_len = p->trail_len;
    e->d_trail.reserve(e->num_props() + e->tail_limit + 64, e->stream);
    if (e->trail_len) CUDA_CHECK(cudaMemcpyAsync(e->d_trail.p, p->d_trail.p, (size_t)e->trail_len * 4, cudaMemcpyDeviceToDevice, e->stream));
    CUDA_CHECK(cudaMalloc(&e->d_ctl, sizeof(Control)));
    Control c;
    std::memset(&c, 0, sizeof(c));
    c.epoch = 1;
    c.trail_cnt = e->trail_len;
    CUDA_CHECK(cudaMemcpyAsync(e->d_ctl, &c, sizeof(c), cudaMemcpyHostToDevice, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    e->at_fixpoint = p->at_fixpoint;
    e->mirror_valid = false;
    e->grid_limit = p->grid_limit;
    e->timing = p->timing;
  });
  if (rc != PCP_OK) {
    parent->err = e->err;
    pcp_engine_destroy(e);
    return rc;
  }
  *out = e;
  return PCP_OK;
}

void pcp_engine_destroy(pcp_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (int f = 0; f < 3; ++f) {
    if (!e->shared) { e->fam[f].d_desc.free(); e->fam[f].d_descB.free(); e->fam[f].d_cdesc.free(); }  // (else: views of the shared block)
    e->fam[f].d_tdesc.free(); e->fam[f].d_tdescB.free(); e->fam[f].d_active.free(); e->fam[f].d_stamp.free();
  }
  if (e->shared) { e->d_adj_ptr = DevBuf<int>(); e->d_adj = DevBuf<uint32_t>(); e->shared.reset(); }
  e->d_nary_ptr.free(); e->d_nary_ops.free(); e->d_nary_kind.free(); e->d_nary_active.free();
  e->d_tree_ptr.free(); e->d_tree_nodes.free();
  e->d_adj_ptr.free(); e->d_adj.free(); e->d_sum_ptr.free(); e->d_sum_terms.free();
  e->d_dirty_bits.free(); e->d_seed_list.free(); e->d_trail.free(); e->d_stack.free(); e->d_bits.free();
  if (e->d_ctl) cudaFree(e->d_ctl);
  if (e->burst.d_bc) cudaFree(e->burst.d_bc);
  e->burst.d_branches.free(); e->burst.d_meta.free(); e->burst.d_bmeta.free(); e->burst.d_tstatus.free(); e->burst.d_tdom.free(); e->burst.d_tbits.free();
  if (e->d_block) cudaFree(e->d_block);
  if (e->h_block) cudaFreeHost(e->h_block);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->d_bbatch) cudaFree(e->d_bbatch);
  if (e->h_bbatch) cudaFreeHost(e->h_bbatch);
  if (e->d_batch) cudaFree(e->d_batch);
  if (e->h_batch) cudaFreeHost(e->h_batch);
  if (e->ev_done) cudaEventDestroy(e->ev_done);
  if (e->ev_sleep) cudaEventDestroy(e->ev_sleep);
  if (e->ev_batch0) cudaEventDestroy(e->ev_batch0);
  if (e->ev_batch1) cudaEventDestroy(e->ev_batch1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

const char* pcp_last_error(const pcp_engine* e) { return e ? e->err.c_str() : "null engine"; }

int pcp_set_timing(pcp_engine* e, int32_t enabled) {
  if (!e) return PCP_ERR_INVALID;
  e->timing = enabled != 0;
  return PCP_OK;
}

int pcp_set_grid_limit(pcp_engine* e, int32_t max_ctas) {
  if (!e || max_ctas < 0) return PCP_ERR_INVALID;
  e->grid_limit = max_ctas;
  return PCP_OK;
}

int pcp_stream(pcp_engine* e, void** stream) {
  if (!e || !stream) return PCP_ERR_INVALID;
  *stream = (void*)e->stream;
  return PCP_OK;
}

int pcp_vars_alloc(pcp_engine* e, const int32_t* lo, const int32_t* hi, int32_t n, int32_t* first_idx) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    PCP_REQUIRE(n >= 0 && (n == 0 || (lo && hi)), "bad arguments");
    for (int i = 0; i < n; ++i) PCP_REQUIRE(lo[i] <= hi[i], "alloc of an empty domain");  // variable/store.rs:136
    for (int i = 0; i < n; ++i)
      PCP_REQUIRE(lo[i] > -kBoundLimit && hi[i] < kBoundLimit, "domain bound outside (-2^29, 2^29)");
    PCP_REQUIRE(e->V + (size_t)n < (size_t)kConstVar28, "too many variables");
    if (first_idx) *first_idx = (int32_t)e->V;
    for (int i = 0; i < n; ++i) { e->h_dom_pending.push_back(make_int2(lo[i], hi[i])); e->h_dom_init.push_back(make_int2(lo[i], hi[i])); }
    e->V += (size_t)n;
    e->at_fixpoint = false;
    e->snapshot_valid = false;
  });
}

int pcp_sum_alloc(pcp_engine* e, const pcp_operand* terms, int32_t n, int32_t* sum_id) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    PCP_REQUIRE(n >= 1 && terms, "At least one variable in sum.");
    long long worst = 0;  // largest magnitude the segmented sum can reach (domains only shrink)
    for (int i = 0; i < n; ++i) {
      PCP_REQUIRE(terms[i].var >= -1, "nested sums are not supported");
      check_operand(e, terms[i]);
      long long m = terms[i].off < 0 ? -(long long)terms[i].off : terms[i].off;
      if (terms[i].var >= 0) {
        const int2 d = e->h_dom_init[(size_t)terms[i].var];
        m = std::max(std::llabs((long long)d.x + terms[i].off), std::llabs((long long)d.y + terms[i].off));
      }
      worst += m;
    }
    PCP_REQUIRE(worst < kBoundLimit, "worst-case range of the Sum view outside (-2^29, 2^29)");
    e->sums.emplace_back(terms, terms + n);
    for (int i = 0; i < n; ++i) e->h_sum_terms.push_back(make_int2(terms[i].var, terms[i].off));
    e->h_sum_ptr.push_back((int)e->h_sum_terms.size());
    if (sum_id) *sum_id = (int32_t)(e->sums.size() - 1);
  });
}

int pcp_props_alloc(pcp_engine* e, int32_t kind, const pcp_operand* ops, int32_t n_ops, int64_t n_props, int32_t* first_idx) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    PCP_REQUIRE(n_props >= 0 && n_ops >= 0 && (n_props == 0 || ops), "bad arguments");
    PCP_REQUIRE(e->num_props() + (size_t)n_props < (size_t)kSlotMask, "too many propagators");
    // all-or-nothing: remember the sizes and roll back on a contract violation
    LabelRec mark;
    for (int f = 0; f < 3; ++f) mark.n_fam[f] = e->fam[f].n;
    mark.n_fam[F_NARY] = e->n_nary;
    mark.n_nary_ops = e->h_nary_ops.size();
    mark.n_tree_ints = e->h_tree_nodes.size();
    mark.n_props = e->num_props();
    int old_max_k = e->nary_max_k;
    if (first_idx) *first_idx = (int32_t)e->num_props();
    try {
      if (n_props > 4096) {  // bulk model upload: one allocation (never for single posts: reserve() is exact)
        if (kind == PCP_X_LESS_Y || kind == PCP_X_NEQ_Y || kind == PCP_X_EQ_Y)
          e->fam[F_BIN].desc.reserve(e->fam[F_BIN].desc.size() + (size_t)n_props);
        e->prop_ref.reserve(e->prop_ref.size() + (size_t)n_props);
      }
      for (int64_t p = 0; p < n_props; ++p) append_prop(e, kind, ops + p * n_ops, n_ops);
    } catch (...) {
      truncate_props(e, mark);
      e->nary_max_k = old_max_k;
      throw;
    }
  });
}

int pcp_prop_alloc(pcp_engine* e, int32_t kind, const pcp_operand* ops, int32_t n_ops, int32_t* idx) {
  return pcp_props_alloc(e, kind, ops, n_ops, 1, idx);
}

int pcp_consistency(pcp_engine* e, int32_t* status, pcp_stats* stats) {
  if (!e || !status) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    CUDA_CHECK(cudaSetDevice(e->device));
    run_fixpoint(e, status, stats);
  });
}

int pcp_consistency_batch(pcp_engine* const* engines, int32_t n, int32_t* status, pcp_stats* stats) {
  if (!engines || n < 0 || !status) return PCP_ERR_INVALID;
  for (int i = 0; i < n; ++i)
    if (!engines[i]) return PCP_ERR_INVALID;
  if (n == 0) return PCP_OK;
  // every launch first, then every wait: the engines' persistent grids run side by side.  An
  // engine without a grid limit of its own gets an equal share of the SMs for this call.
  int rc = PCP_OK, launched = 0;
  pcp_engine* lead = engines[0];
  const bool timed = lead->timing;
  std::vector<int> saved((size_t)n);
  for (int i = 0; i < n; ++i) saved[(size_t)i] = engines[i]->grid_limit;
  // host work first (uploads, buffer growth, parameters), then the launches back to back
  int prepared = 0;
  for (int i = 0; i < n && rc == PCP_OK; ++i) {
    pcp_engine* e = engines[i];
    rc = guarded(e, [&] {
      PCP_REQUIRE_NO_BURST(e);
      CUDA_CHECK(cudaSetDevice(e->device));
      if (e->grid_limit == 0 && n > 1) e->grid_limit = std::max(1, e->num_sms / n);
      fixpoint_prepare(e);
    });
    if (rc == PCP_OK) ++prepared;
  }
  // (an engine whose preparation failed is left out; the ones already prepared have consumed their
  // pending restore / uploads and must run, so the call goes on for them and reports the error)
  const int first_rc = rc;
  rc = PCP_OK;
  bool fused = false;
  if (prepared == n && batch_fusable(engines, n)) {
    rc = guarded(lead, [&] {
      CUDA_CHECK(cudaSetDevice(lead->device));
      fire_fused(engines, n, timed);
    });
    fused = rc == PCP_OK;
    if (fused) launched = n;
  }
  for (int i = 0; i < prepared && rc == PCP_OK && !fused; ++i) {
    pcp_engine* e = engines[i];
    rc = guarded(e, [&] {
      CUDA_CHECK(cudaSetDevice(e->device));
      if (timed && i == 0) CUDA_CHECK(cudaEventRecord(lead->ev_batch0, lead->stream));
      if (timed && i > 0) CUDA_CHECK(cudaStreamWaitEvent(e->stream, lead->ev_batch0, 0));
      fixpoint_fire(e);
      if (timed && i > 0) {
        CUDA_CHECK(cudaEventRecord(e->ev_done, e->stream));
        CUDA_CHECK(cudaStreamWaitEvent(lead->stream, e->ev_done, 0));
      }
    });
    if (rc == PCP_OK) ++launched;
  }
  if (timed && launched == n && !fused) {
    int r2 = guarded(lead, [&] { CUDA_CHECK(cudaSetDevice(lead->device)); CUDA_CHECK(cudaEventRecord(lead->ev_batch1, lead->stream)); });
    if (rc == PCP_OK) rc = r2;
  }
  for (int i = 0; i < launched; ++i) {
    pcp_engine* e = engines[i];
    int r2 = guarded(e, [&] {
      CUDA_CHECK(cudaSetDevice(e->device));
      fixpoint_wait(e, &status[i], stats ? &stats[i] : nullptr);
    });
    if (rc == PCP_OK) rc = r2;
  }
  for (int i = 0; i < n; ++i) engines[i]->grid_limit = saved[(size_t)i];
  if (first_rc != PCP_OK) return first_rc;
  if (timed && launched == n && rc == PCP_OK && stats) {
    rc = guarded(lead, [&] {
      CUDA_CHECK(cudaEventSynchronize(lead->ev_batch1));
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, lead->ev_batch0, lead->ev_batch1));
      stats[0].kernel_ms = ms;  // the whole batch: first launch to last completion
    });
  }
  if (fused && stats) stats[0].launches = 1;
  return rc;
}

int pcp_domains_read(pcp_engine* e, int32_t first, int32_t n, int32_t* lo, int32_t* hi) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE(first >= 0 && n >= 0 && (size_t)first + (size_t)n <= e->V, "Variable not registered in the store.");
    if (n == 0) return;
    if (!e->mirror_valid || prologue_pending(e)) {
      CUDA_CHECK(cudaSetDevice(e->device));
      fetch_domains(e);
    }
    const int2* d = e->h_dom() + first;
    for (int i = 0; i < n; ++i) { lo[i] = d[i].x; hi[i] = d[i].y; }
  });
}

int pcp_domains_size_read(pcp_engine* e, int32_t first, int32_t n, uint32_t* size) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE(first >= 0 && n >= 0 && (size_t)first + (size_t)n <= e->V && (n == 0 || size), "Variable not registered in the store.");
    if (n == 0) return;
    if (!e->mirror_valid || prologue_pending(e)) {
      CUDA_CHECK(cudaSetDevice(e->device));
      fetch_domains(e);
      e->sizes_valid = false;
    }
    if (!e->set_mode) {  // Interval::size()
      const int2* d = e->h_dom() + first;
      for (int i = 0; i < n; ++i) size[i] = d[i].x > d[i].y ? 0u : (uint32_t)(d[i].y - d[i].x) + 1u;
      return;
    }
    if (!e->sizes_valid) {
      CUDA_CHECK(cudaSetDevice(e->device));
      pcp_set_sizes_kernel<<<(int)((e->V + 255) / 256), 256, 0, e->stream>>>(e->d_dom(), e->d_bits.p, e->set_W, e->set_base, (int)e->V,
                                                                             e->h_sizes_dev());
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaStreamSynchronize(e->stream));
      e->sizes_valid = true;
    }
    std::memcpy(size, e->h_sizes() + first, (size_t)n * sizeof(uint32_t));
  });
}

int pcp_domains_read_bits(pcp_engine* e, int32_t first, int32_t n, int32_t base, int32_t words, uint32_t* out) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE(first >= 0 && n >= 0 && words >= 0 && (size_t)first + (size_t)n <= e->V && (n == 0 || words == 0 || out),
                "Variable not registered in the store.");
    if (n == 0 || words == 0) return;
    if (!e->mirror_valid || prologue_pending(e)) {
      CUDA_CHECK(cudaSetDevice(e->device));
      fetch_domains(e);
    }
    std::vector<uint32_t> raw;
    if (e->set_mode) {
      raw.resize((size_t)n * e->set_W);
      CUDA_CHECK(cudaSetDevice(e->device));
      CUDA_CHECK(cudaMemcpyAsync(raw.data(), e->d_bits.p + (size_t)first * e->set_W, raw.size() * 4, cudaMemcpyDeviceToHost, e->stream));
      CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    std::memset(out, 0, (size_t)n * (size_t)words * 4);
    for (int i = 0; i < n; ++i) {
      const int2 d = e->h_dom()[first + i];
      for (long long v = d.x; v <= d.y; ++v) {
        if (e->set_mode) {
          const unsigned b = (unsigned)(v - e->set_base);
          if (!((raw[(size_t)i * e->set_W + (b >> 5)] >> (b & 31)) & 1u)) continue;
        }
        const long long o = v - base;
        PCP_REQUIRE(o >= 0 && o < (long long)words * 32, "value outside the requested bit window");
        out[(size_t)i * words + (size_t)(o >> 5)] |= 1u << (o & 31);
      }
    }
  });
}

static_assert(PCP_F_MAX_NODES == kTreeMaxNodes && PCP_F_MAX_VARS == kTreeMaxVars, "header limits out of step with the device");
namespace {
// Host form of a formula tree (pcp_formula_alloc): parsed from the prefix words, negations pushed to
// the leaves the way the reference's NotFormula impls do, then flattened for the device.
struct FTree {
  int type = 0;                 // TN_CONJ, TN_DISJ, TN_BOOL, TN_BOOL_NEG, TN_LEAF + BinKind, TN_LEAF3 + TerKind
  std::vector<FTree> kids;
  pcp_operand op[3];
  int nops = 0;
};

FTree ftree_negate(const FTree& f) {
  FTree g;
  switch (f.type) {
    case TN_CONJ: case TN_DISJ:  // De Morgan (conjunction.rs:66-74, disjunction.rs:66-74)
      g.type = f.type == TN_CONJ ? TN_DISJ : TN_CONJ;
      for (const FTree& k : f.kids) g.kids.push_back(ftree_negate(k));
      return g;
    case TN_BOOL: g = f; g.type = TN_BOOL_NEG; return g;       // boolean.rs:70-79
    case TN_BOOL_NEG: g = f; g.type = TN_BOOL; return g;       // boolean_neg.rs:58-66
    case TN_LEAF + B_LESS:   // not (x < y) = x >= y = x_greater_y(x + 1, y) = XLessY(y, x + 1)  (x_less_y.rs:57-65, cmp/mod.rs:40-52)
      g.type = TN_LEAF + B_LESS; g.nops = 2; g.op[0] = f.op[1]; g.op[1] = f.op[0]; g.op[1].off += 1; return g;
    case TN_LEAF + B_NEQ: g = f; g.type = TN_LEAF + B_EQ; return g;    // x_neq_y.rs:56-64
    case TN_LEAF + B_EQ: g = f; g.type = TN_LEAF + B_NEQ; return g;    // x_eq_y.rs:57-65
    case TN_LEAF3 + T_GREATER:  // not (x > y + z) = x <= y + z = XLessYPlusZ(x - 1, y, z)  (cmp/mod.rs:78-86)
      g = f; g.type = TN_LEAF3 + T_LESS; g.op[0].off -= 1; return g;
    case TN_LEAF3 + T_LESS:     // not (x < y + z) = x >= y + z = XGreaterYPlusZ(x + 1, y, z)  (cmp/mod.rs:70-76)
      g = f; g.type = TN_LEAF3 + T_GREATER; g.op[0].off += 1; return g;
    default:                    // XEqYPlusZ::not is unimplemented!() in the reference (x_eq_y_plus_z.rs:70-77)
      PCP_FAIL(PCP_ERR_UNSUPPORTED, "not() of XEqYPlusZ is unimplemented in the reference");
  }
}

FTree ftree_parse(pcp_engine* e, const int32_t*& p, const int32_t* end, int depth) {
  PCP_REQUIRE(p < end, "truncated formula");
  PCP_REQUIRE(depth < 16, "formula nested too deeply");
  const int32_t tag = *p++;
  FTree f;
  auto operand = [&](int i) {
    PCP_REQUIRE(p + 2 <= end, "truncated formula");
    pcp_operand raw{p[0], p[1]};
    p += 2;
    f.op[i] = lower_view(e, raw);
    if (f.op[i].var <= -2) PCP_FAIL(PCP_ERR_UNSUPPORTED, "multi-term Sum views inside a formula tree have no device lowering");
  };
  if (tag == PCP_F_CONJUNCTION || tag == PCP_F_DISJUNCTION) {
    PCP_REQUIRE(p < end && *p >= 1, "connective without children");
    const int32_t n = *p++;
    f.type = tag == PCP_F_CONJUNCTION ? TN_CONJ : TN_DISJ;
    for (int32_t i = 0; i < n; ++i) f.kids.push_back(ftree_parse(e, p, end, depth + 1));
    return f;
  }
  if (tag == PCP_F_NOT) return ftree_negate(ftree_parse(e, p, end, depth + 1));
  if (tag == PCP_F_BOOLEAN || tag == PCP_F_BOOLEAN_NEG) {
    f.type = tag == PCP_F_BOOLEAN ? TN_BOOL : TN_BOOL_NEG;
    f.nops = 1;
    operand(0);
    PCP_REQUIRE(f.op[0].var >= 0 && f.op[0].off == 0, "Boolean reads a variable");
    const int2 d = e->h_dom_init[(size_t)f.op[0].var];
    PCP_REQUIRE(d.x >= 0 && d.y <= 1, "Boolean variables have the domain [0, 1] (boolean.rs:35-40)");
    return f;
  }
  const int kind = tag - PCP_F_LEAF;
  switch (kind) {
    case PCP_X_LESS_Y: f.type = TN_LEAF + B_LESS; f.nops = 2; break;
    case PCP_X_NEQ_Y: f.type = TN_LEAF + B_NEQ; f.nops = 2; break;
    case PCP_X_EQ_Y: f.type = TN_LEAF + B_EQ; f.nops = 2; break;
    case PCP_X_GREATER_Y_PLUS_Z: f.type = TN_LEAF3 + T_GREATER; f.nops = 3; break;
    case PCP_X_LESS_Y_PLUS_Z: f.type = TN_LEAF3 + T_LESS; f.nops = 3; break;
    case PCP_X_EQ_Y_PLUS_Z: f.type = TN_LEAF3 + T_EQ; f.nops = 3; break;
    default: PCP_FAIL(PCP_ERR_UNSUPPORTED, "unknown formula node / leaf kind without a device lowering");
  }
  for (int i = 0; i < f.nops; ++i) operand(i);
  require_distinct_vars(e, f.op, f.nops);  // a leaf subscribing twice to one variable (indexed_deps.rs:69-77)
  return f;
}

// Conjunction / Disjunction::dependencies sort and de-duplicate (variable, event) pairs
// (conjunction.rs:107-118, disjunction.rs:118-129), so one variable may be read by several leaves
// -- but only at one event: XLessY / the ternaries / Boolean subscribe at Bound, XNeqY / XEqY at
// Inner, and a propagator that ends up on two lists of one variable trips the assert of
// IndexedDeps::subscribe (indexed_deps.rs:69-77) when the store prepares.  ev[var] = bit set of events.
void ftree_events(const FTree& f, std::vector<std::pair<int, int>>& ev) {
  for (const FTree& k : f.kids) ftree_events(k, ev);
  if (f.type == TN_CONJ || f.type == TN_DISJ) return;
  const bool inner = f.type == TN_LEAF + B_NEQ || f.type == TN_LEAF + B_EQ;
  for (int i = 0; i < f.nops; ++i)
    if (f.op[i].var >= 0) ev.emplace_back(f.op[i].var, inner ? 2 : 1);
}

int ftree_count(const FTree& f) {
  int n = 1;
  for (const FTree& k : f.kids) n += ftree_count(k);
  return n;
}
// prefix order; node indices are relative to the tree's first node
void ftree_flatten(const FTree& f, std::vector<int>& out, std::vector<int>& vars, int base_node) {
  const size_t at = out.size();
  out.resize(at + kTreeNodeInts, 0);
  out[at + 0] = f.type;
  out[at + 1] = (int)f.kids.size();
  for (int i = 0; i < 3; ++i) {
    int slot = -1, off = 0;
    if (i < f.nops) {
      off = f.op[i].off;
      if (f.op[i].var >= 0) {
        auto it = std::find(vars.begin(), vars.end(), f.op[i].var);
        if (it == vars.end()) { vars.push_back(f.op[i].var); it = vars.end() - 1; }
        slot = (int)(it - vars.begin());
      }
    }
    out[at + 3 + 2 * i] = slot;
    out[at + 4 + 2 * i] = off;
  }
  for (const FTree& k : f.kids) ftree_flatten(k, out, vars, base_node);
  out[at + 2] = (int)((out.size() - (size_t)base_node) / kTreeNodeInts);
}
}  // namespace

int pcp_formula_alloc(pcp_engine* e, const int32_t* words, int32_t n_words, int32_t* idx) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    PCP_REQUIRE(words && n_words > 0, "bad arguments");
    if (e->set_mode) PCP_FAIL(PCP_ERR_UNSUPPORTED, "formula trees have no device lowering on IntervalSet domains");
    const int32_t* p = words;
    const FTree f = ftree_parse(e, p, words + n_words, 0);
    PCP_REQUIRE(p == words + n_words, "trailing words after the formula");
    {
      std::vector<std::pair<int, int>> ev;
      ftree_events(f, ev);
      std::sort(ev.begin(), ev.end());
      for (size_t i = 1; i < ev.size(); ++i)
        PCP_REQUIRE(ev[i].first != ev[i - 1].first || ev[i].second == ev[i - 1].second, "propagator already subscribed to this variable");
    }
    if (ftree_count(f) > kTreeMaxNodes) PCP_FAIL(PCP_ERR_UNSUPPORTED, "formula tree with more than PCP_F_MAX_NODES nodes");
    std::vector<int> nodes, vars;
    ftree_flatten(f, nodes, vars, 0);
    if ((int)vars.size() > kTreeMaxVars) PCP_FAIL(PCP_ERR_UNSUPPORTED, "formula tree over more than PCP_F_MAX_VARS variables");
    if (idx) *idx = (int32_t)e->num_props();
    e->h_tree_ptr.push_back((int)e->h_tree_nodes.size());
    e->h_tree_nodes.insert(e->h_tree_nodes.end(), nodes.begin(), nodes.end());
    for (int v : vars) e->h_nary_ops.push_back(make_int2(v, 0));  // the variable table = the n-ary operands
    e->h_nary_ptr.push_back((int)e->h_nary_ops.size());
    e->h_nary_kind.push_back((int)N_TREE);
    e->nary_max_k = std::max(e->nary_max_k, (int)vars.size());
    e->prop_ref.push_back(make_ref(F_NARY, (unsigned)e->n_nary++));
    e->at_fixpoint = false;  // a new n-ary propagator needs a full first sweep
  });
}

int pcp_var_update(pcp_engine* e, int32_t idx, int32_t lo, int32_t hi, int32_t* ok) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    PCP_REQUIRE(idx >= 0 && (size_t)idx < e->V, "Variable not registered in the store.");
    CUDA_CHECK(cudaSetDevice(e->device));
    fetch_domains(e);
    int2 cur = e->h_dom()[idx];
    bool empty = lo > hi;
    // variable/store.rs:153-156: dom must be a subset of the current domain (empty is)
    PCP_REQUIRE(empty || (lo >= cur.x && hi <= cur.y), "Domain update must be monotonic.");
    if (!empty && e->set_mode) {
      // IntervalSet: the new domain is cur /\ [lo, hi]; its bounds are the extreme values left
      std::vector<uint32_t> w((size_t)e->set_W);
      CUDA_CHECK(cudaMemcpyAsync(w.data(), e->d_bits.p + (size_t)idx * e->set_W, w.size() * 4, cudaMemcpyDeviceToHost, e->stream));
      CUDA_CHECK(cudaStreamSynchronize(e->stream));
      auto has = [&](int v) { unsigned b = (unsigned)(v - e->set_base); return (w[b >> 5] >> (b & 31)) & 1u; };
      while (lo <= hi && !has(lo)) ++lo;
      while (hi >= lo && !has(hi)) --hi;
      empty = lo > hi;
    }
    if (empty) { if (ok) *ok = 0; return; }
    if (lo != cur.x || hi != cur.y) {
      int2* st = static_cast<int2*>(stage(e, sizeof(int2)));
      *st = make_int2(lo, hi);
      CUDA_CHECK(cudaMemcpyAsync(e->d_dom() + idx, st, sizeof(int2), cudaMemcpyHostToDevice, e->stream));
      CUDA_CHECK(cudaStreamSynchronize(e->stream));
      e->h_dom()[idx] = make_int2(lo, hi);
      e->sizes_valid = false;
      e->host_dirty.push_back(idx);
      ++e->dom_version;
      e->snapshot_valid = false;
    }
    if (ok) *ok = 1;
  });
}

int pcp_active_read(pcp_engine* e, int32_t first, int32_t n, uint8_t* out) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE(first >= 0 && n >= 0 && (size_t)first + (size_t)n <= e->num_props(), "propagator out of range");
    if (n == 0) return;
    CUDA_CHECK(cudaSetDevice(e->device));
    sync_device_state(e);
    std::vector<uint32_t> bits[4];
    for (int f = 0; f < 4; ++f) {
      size_t cnt = f < 3 ? e->fam[f].n : e->n_nary;
      const uint32_t* src = f < 3 ? e->fam[f].d_active.p : e->d_nary_active.p;
      bits[f].resize((cnt + 31) / 32);
      if (!bits[f].empty())
        CUDA_CHECK(cudaMemcpyAsync(bits[f].data(), src, bits[f].size() * 4, cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < n; ++i) {
      uint32_t ref = e->prop_ref[(size_t)first + i];
      uint32_t slot = ref & kSlotMask;
      out[i] = (bits[ref >> 29][slot >> 5] >> (slot & 31)) & 1u;
    }
  });
}

int pcp_label(pcp_engine* e, uint64_t* label) {
  if (!e || !label) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    PCP_REQUIRE(e->labels.size() < e->max_labels, "label stack full (pcp_config.max_labels)");
    size_t idx = e->labels.size();
    const bool have = e->snapshot_valid && e->snapshot_version == e->dom_version && !prologue_pending(e);
    if (!have) {
      CUDA_CHECK(cudaSetDevice(e->device));
      sync_device_state(e);
      if (e->stack_stride != e->slot_stride()) { PCP_REQUIRE(e->labels.empty(), "variables allocated while labels are live"); e->stack_stride = e->slot_stride(); }
      if ((idx + 1) * std::max<size_t>(e->stack_stride, 1) > e->d_stack.cap)  // grow in big steps: a reallocation copies the stack
        e->d_stack.reserve(std::max<size_t>((idx + 1) * 2, std::min<size_t>(e->max_labels, 256)) * std::max<size_t>(e->stack_stride, 1),
                           e->stream, idx * e->stack_stride);
      if (e->V)
        CUDA_CHECK(cudaMemcpyAsync(e->d_stack.p + idx * e->stack_stride, e->d_dom(), e->V * sizeof(int2),
                                   cudaMemcpyDeviceToDevice, e->stream));
      if (e->V && e->set_mode)
        CUDA_CHECK(cudaMemcpyAsync(e->d_stack.p + idx * e->stack_stride + e->V, e->d_bits.p,
                                   e->V * (size_t)e->set_W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
    }
    e->snapshot_valid = false;
    LabelRec r;
    for (int f = 0; f < 3; ++f) r.n_fam[f] = e->fam[f].n;
    r.n_fam[F_NARY] = e->n_nary;
    r.n_nary_ops = e->h_nary_ops.size();
    r.n_tree_ints = e->h_tree_nodes.size();
    r.n_props = e->num_props();
    r.trail_len = e->trail_len;
    r.at_fixpoint = e->at_fixpoint;
    r.dom_version = e->dom_version;
    e->labels.push_back(r);
    *label = idx;
  });
}

int pcp_restore(pcp_engine* e, uint64_t label) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    PCP_REQUIRE(label < e->labels.size(), "unknown label (invalidated by an earlier restore?)");
    const LabelRec r = e->labels[label];
    e->labels.resize(label + 1);
    truncate_props(e, r);
    e->snapshot_valid = false;
    e->sizes_valid = false;
    e->host_dirty.clear();
    e->at_fixpoint = r.at_fixpoint;
    // restoring the label that was just taken from the current domains (the left child of
    // Branch::distribute) needs no copy
    const bool same_domains = r.dom_version == e->dom_version && !e->pending_restore;
    if (!same_domains) {
      e->pending_restore = true;
      e->pending_restore_label = label;
      e->mirror_valid = false;
    }
    e->pending_trail_undo = true;
    e->pending_trail_keep = r.trail_len;
  });
}

// ---------------------------------------------------------------------------------------
// device-resident search bursts (private interface, pcp_internal.h)
// ---------------------------------------------------------------------------------------
namespace {
constexpr size_t kBurstStackBudget = (size_t)8 << 30;  // bytes of label slots a device search may reserve
size_t burst_depth(const pcp_engine* e) { return std::max<size_t>(e->max_labels, 16384); }
}  // namespace

int pcp_internal_burst_supported(pcp_engine* e, const pcp_search_config* cfg, uint64_t trace_capacity) {
  if (!e || !cfg) return 0;
  static const bool off = std::getenv("PCP_NO_BURST") != nullptr;
  if (off) return 0;
  if (cfg->var_sel != 0 || cfg->val_sel != 0 || cfg->distributor != 0 || cfg->bb_mode != 0) return 0;
  if (e->V == 0 || e->V > (size_t)(1 << 20)) return 0;
  if ((e->flags & PCP_FLAG_HOST_SEARCH)) return 0;
  // the device search pre-reserves its whole label stack (burst_depth(e) slots of V domains):
  // a store too wide for that budget takes the host-driven node loop, which grows on demand
  if (burst_depth(e) * e->slot_stride() * sizeof(int2) > kBurstStackBudget) return 0;
  // the device trace keeps full domains for the traced nodes
  if (trace_capacity * e->slot_stride() * sizeof(int2) > (size_t)1 << 30) return 0;
  return 1;
}

int pcp_internal_burst_begin(pcp_engine* e, int32_t all_solutions, uint64_t node_limit, uint64_t trace_capacity,
                             int32_t trace_domains) {
  if (!e) return PCP_ERR_INVALID;
  if (e->burst.open) { e->err = "a device search is already open on this engine"; return PCP_ERR_INVALID; }
  int rc = pcp_label(e, &e->burst.root_label);  // flushes everything pending; the root of the search
  if (rc != PCP_OK) return rc;
  rc = guarded(e, [&] {
    CUDA_CHECK(cudaSetDevice(e->device));
    auto& b = e->burst;
    const size_t V = e->V;
    // room for the search: label slots, branching constraints in the binary tail, trail
    const size_t depth = burst_depth(e);
    b.max_labels = (int)(e->labels.size() + depth);
    e->stack_stride = e->slot_stride();
    e->d_stack.reserve((size_t)b.max_labels * e->stack_stride, e->stream, e->labels.size() * e->stack_stride);
    HostFamily& hb = e->fam[F_BIN];
    b.bin_cap = (int)(hb.n + depth);
    if (e->shared) {
      const size_t t0 = hb.tail_base;
      if ((size_t)b.bin_cap - t0 > hb.d_tdesc.cap) hb.d_tdesc.reserve((size_t)b.bin_cap - t0, e->stream, hb.n - t0);
    } else {
      hb.d_desc.reserve((size_t)b.bin_cap, e->stream, hb.n);
    }
    reserve_zeroed(e, hb.d_active, ((size_t)b.bin_cap + 31) / 32 + 1, (hb.active_set + 31) / 32 + 1);
    if ((size_t)b.bin_cap > hb.d_stamp.cap) {
      size_t valid = std::min(hb.d_stamp.cap, hb.n);
      hb.d_stamp.reserve((size_t)b.bin_cap, e->stream, valid);
      fill_u32(e, hb.d_stamp.p + valid, 0u, hb.d_stamp.cap - valid);
    }
    e->d_trail.reserve(e->num_props() + depth + 64, e->stream, e->trail_len);
    b.d_branches.reserve(2 * depth + 16, e->stream);
    b.d_bmeta.reserve(2 * depth + 16, e->stream);
    b.d_meta.reserve((size_t)b.max_labels, e->stream);
    b.trace_cap = trace_capacity;
    b.trace_dom = trace_domains != 0 || trace_capacity > 0;
    if (trace_capacity) {
      b.d_tstatus.reserve(trace_capacity, e->stream);
      if (b.trace_dom) b.d_tdom.reserve(trace_capacity * V, e->stream);
      if (b.trace_dom && e->set_mode) b.d_tbits.reserve(trace_capacity * V * (size_t)e->set_W, e->stream);
    }
    if (!b.d_bc) CUDA_CHECK(cudaMalloc(&b.d_bc, sizeof(BurstCtl)));
    BurstCtl bc;
    std::memset(&bc, 0, sizeof(bc));
    bc.cmd = 1;
    bc.root_pending = 1;
    bc.inl_slot = -1;
    bc.bin_n = (int)hb.n;
    bc.n_labels = (int)e->labels.size();
    bc.cur_label = -1;
    CUDA_CHECK(cudaMemcpyAsync(b.d_bc, &bc, sizeof(bc), cudaMemcpyHostToDevice, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    b.P = prepare(e);  // nothing pending: pure parameter block
    b.P.full_sweep = 1;
    b.P.sync0 = 0;
    b.P.do_trail = 0;
    b.P.restore_from = nullptr;
    b.P.n_inline = 0;
    b.P.snapshot_to = nullptr;
    b.P.trace = nullptr;
    b.all_solutions = all_solutions;
    b.node_limit = node_limit;
    b.props_base = (long long)e->num_props() - (long long)hb.n;
    b.props0 = e->props_total;
    b.kernel_seconds = 0;
    b.open = true;
  });
  if (rc != PCP_OK) e->labels.resize((size_t)e->burst.root_label);  // give the root label back
  return rc;
}

namespace {
struct BurstLaunch {
  BurstParams B;
  const void* fn = nullptr;
  int grid = 0;
  size_t smem = 0;
  bool smem_dom = false;
};

// launch parameters of the next slice of e's device search (e->burst.P is completed in place)
void burst_prepare(pcp_engine* e, uint64_t max_nodes, BurstLaunch& L) {
  auto& b = e->burst;
  PCP_REQUIRE(b.open, "no device search is open");
  const size_t V = e->V;
  BurstParams& B = L.B;
  std::memset(&B, 0, sizeof(B));
  B.bc = b.d_bc;
  B.branches = b.d_branches.p;
  B.label_meta = b.d_meta.p;
  B.branch_meta = b.d_bmeta.p;
  B.stack = e->d_stack.p;
  B.stack_stride = (long long)e->stack_stride;
  B.max_labels = b.max_labels;
  B.max_branches = (int)b.d_branches.cap;
  B.bin_cap = b.bin_cap;
  B.all_solutions = b.all_solutions;
  B.incremental = (e->flags & PCP_FLAG_INCREMENTAL) ? 1 : 0;
  B.node_budget = max_nodes ? max_nodes : ~0ull;
  B.node_limit = b.node_limit;
  B.props_base = b.props_base;
  B.t_status = b.trace_cap ? b.d_tstatus.p : nullptr;
  B.t_dom = (b.trace_cap && b.trace_dom) ? b.d_tdom.p : nullptr;
  B.t_bits = (b.trace_cap && b.trace_dom && e->set_mode) ? b.d_tbits.p : nullptr;
  B.t_cap = b.trace_cap;
  Params& P = b.P;
  P.max_iterations = e->max_iterations;
  if (e->epoch > 0x70000000u) {  // epoch wrap, as in run_fixpoint
    for (int f = 0; f < 3; ++f) fill_u32(e, e->fam[f].d_stamp.p, 0u, e->fam[f].d_stamp.cap);
    e->epoch = 1;
  }
  P.gen0 = e->bar_gen;
  P.epoch0 = e->epoch;
  size_t total = e->n_nary * 4096;
  for (int f = 0; f < 3; ++f) total += e->fam[f].n;
  L.grid = (int)std::min<size_t>((size_t)launch_ctas(e), std::max<size_t>(1, (total + 4095) / 4096));
  size_t dom_bytes = (V * 8 + 15) & ~size_t(15);
  L.smem_dom = (size_t)kRingBytes + dom_bytes + e->static_smem + 256 <= (size_t)e->max_smem_optin;
  L.smem = (size_t)kRingBytes + (L.smem_dom ? dom_bytes : 0);
  P.smem_dom = L.smem_dom ? 1 : 0;
  {
    const size_t bm_bytes = (e->dirty_words * 4 + 15) & ~size_t(15);
    P.dirty_bm_off = 0;
    if (!L.smem_dom && L.smem + bm_bytes + e->static_smem + 256 <= (size_t)e->max_smem_optin) {
      P.dirty_bm_off = (int)L.smem;
      L.smem += bm_bytes;
    }
  }
  L.fn = e->set_mode ? burst_set_fn(bin_only_store(e), L.smem_dom) : burst_fn(bin_only_store(e), L.smem_dom);
}

// host bookkeeping after a slice: `bc` and the result header have been copied back
void burst_finish(pcp_engine* e, const BurstCtl& bc, pcp_burst_result* res) {
  auto& b = e->burst;
  const Result& r = *e->h_result();
  e->epoch = r.epoch;
  e->bar_gen = r.gen;
  e->trail_len = r.trail_cnt;
  e->props_total = r.propagations;
  e->mirror_valid = false;
  e->snapshot_valid = false;
  e->at_fixpoint = false;
  ++e->dom_version;
  if (bc.err == 2) PCP_FAIL(PCP_ERR_CUDA, "fixpoint iteration cap reached");
  if (bc.err == 3) PCP_FAIL(PCP_ERR_CUDA, "device search: the bookkeeping CTA and the grid derived different branching decisions");
  if (bc.err) PCP_FAIL(PCP_ERR_NOMEM, "device search: label / branch / tail capacity exceeded (pcp_config.max_labels)");
  res->status = bc.status;
  res->err = bc.err;
  res->nodes = bc.nodes;
  res->solutions = bc.solutions;
  res->failures = bc.failures;
  res->iterations = bc.iterations;
  res->propagations = r.propagations - b.props0;
  res->kernel_seconds = b.kernel_seconds;
}
}  // namespace

int pcp_internal_burst_step(pcp_engine* e, uint64_t max_nodes, pcp_burst_result* res) {
  if (!e || !res) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    auto& b = e->burst;
    CUDA_CHECK(cudaSetDevice(e->device));
    BurstLaunch L;
    burst_prepare(e, max_nodes, L);
    CUDA_CHECK(cudaEventRecord(e->ev0, e->stream));
    void* args[] = {&b.P, &L.B};
    launch_persistent(L.fn, L.grid, args, L.smem, e->stream);
    CUDA_CHECK(cudaEventRecord(e->ev1, e->stream));
    BurstCtl bc;
    CUDA_CHECK(cudaMemcpyAsync(&bc, b.d_bc, sizeof(bc), cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaMemcpyAsync(e->h_block, e->d_block, sizeof(Result), cudaMemcpyDeviceToHost, e->stream));
    // (a blocking wait: with many searches side by side -- one host thread each -- spinning threads
    // would outnumber the cores)
    CUDA_CHECK(cudaEventRecord(e->ev_sleep, e->stream));
    CUDA_CHECK(cudaEventSynchronize(e->ev_sleep));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    b.kernel_seconds += ms * 1e-3;
    burst_finish(e, bc, res);
  });
}

// The next slice of several device searches in ONE launch (pcp_burst_batch_kernel: a group of CTAs
// per search).  *fused = 0 and nothing done when the engines cannot share a launch (different kernel
// variants or geometries, no shared-memory snapshot, more CTAs than SMs): the caller then steps them
// one by one.  budgets[i] = node budget of search i for this slice.
int pcp_internal_burst_step_many(pcp_engine* const* es, int32_t n, const uint64_t* budgets, pcp_burst_result* res, int32_t* fused) {
  if (!es || n <= 0 || !budgets || !res || !fused) return PCP_ERR_INVALID;
  *fused = 0;
  pcp_engine* lead = es[0];
  std::vector<BurstLaunch> Ls((size_t)n);
  int rc = PCP_OK;
  {
    const char* sw = std::getenv("PCP_BATCH");  // measurement / test switch
    if (sw && std::strcmp(sw, "unfused") == 0) return PCP_OK;
  }
  // can they share a launch?  (decided before anything is changed)
  for (int i = 0; i < n; ++i) {
    pcp_engine* e = es[i];
    if (!e || !e->burst.open || e->device != lead->device || e->set_mode != lead->set_mode ||
        bin_only_store(e) != bin_only_store(lead))
      return PCP_OK;
    for (int j = 0; j < i; ++j)
      if (es[j] == e) return PCP_OK;
  }
  for (int i = 0; i < n && rc == PCP_OK; ++i) {
    pcp_engine* e = es[i];
    rc = guarded(e, [&] {
      CUDA_CHECK(cudaSetDevice(e->device));
      burst_prepare(e, budgets[i], Ls[(size_t)i]);  // (idempotent: nothing is consumed before the launch)
    });
  }
  if (rc != PCP_OK) return rc;
  for (int i = 0; i < n; ++i)
    if (!Ls[(size_t)i].smem_dom || Ls[(size_t)i].grid != Ls[0].grid || Ls[(size_t)i].smem != Ls[0].smem) return PCP_OK;
  if ((long long)n * Ls[0].grid > lead->num_sms) return PCP_OK;
  rc = guarded(lead, [&] {
    const size_t bytes_p = (size_t)n * sizeof(Params), bytes_b = (size_t)n * sizeof(BurstParams);
    if (lead->burst_batch_cap < n) {
      if (lead->d_bbatch) cudaFree(lead->d_bbatch);
      if (lead->h_bbatch) cudaFreeHost(lead->h_bbatch);
      lead->d_bbatch = lead->h_bbatch = nullptr;
      lead->burst_batch_cap = 0;
      const size_t cap = (size_t)std::max(n, 32);
      CUDA_CHECK(cudaMalloc(&lead->d_bbatch, cap * (sizeof(Params) + sizeof(BurstParams)) + cap * (sizeof(BurstCtl) + sizeof(Result))));
      CUDA_CHECK(cudaHostAlloc(&lead->h_bbatch, cap * (sizeof(Params) + sizeof(BurstParams)) + cap * (sizeof(BurstCtl) + sizeof(Result)), cudaHostAllocDefault));
      lead->burst_batch_cap = (int)cap;
    }
    const size_t cap = (size_t)lead->burst_batch_cap;
    char* hb = static_cast<char*>(lead->h_bbatch);
    char* db = static_cast<char*>(lead->d_bbatch);
    Params* hP = reinterpret_cast<Params*>(hb);
    BurstParams* hB = reinterpret_cast<BurstParams*>(hb + cap * sizeof(Params));
    const Params* dP = reinterpret_cast<const Params*>(db);
    const BurstParams* dB = reinterpret_cast<const BurstParams*>(db + cap * sizeof(Params));
    BurstCtl* hbc = reinterpret_cast<BurstCtl*>(hb + cap * (sizeof(Params) + sizeof(BurstParams)));
    for (int i = 0; i < n; ++i) { hP[i] = es[i]->burst.P; hB[i] = Ls[(size_t)i].B; }
    for (int i = 1; i < n; ++i) {
      CUDA_CHECK(cudaEventRecord(es[i]->ev_done, es[i]->stream));
      CUDA_CHECK(cudaStreamWaitEvent(lead->stream, es[i]->ev_done, 0));
    }
    CUDA_CHECK(cudaMemcpyAsync(db, hb, bytes_p, cudaMemcpyHostToDevice, lead->stream));
    CUDA_CHECK(cudaMemcpyAsync(db + cap * sizeof(Params), hb + cap * sizeof(Params), bytes_b, cudaMemcpyHostToDevice, lead->stream));
    const void* fn = lead->set_mode ? (bin_only_store(lead) ? (const void*)binonly_set::pcp_burst_batch_kernel : (const void*)full_set::pcp_burst_batch_kernel)
                                    : (bin_only_store(lead) ? (const void*)binonly::pcp_burst_batch_kernel : (const void*)full::pcp_burst_batch_kernel);
    int group = Ls[0].grid;
    void* args[] = {(void*)&dP, (void*)&dB, (void*)&group};
    CUDA_CHECK(cudaEventRecord(lead->ev0, lead->stream));
    launch_persistent(fn, n * group, args, Ls[0].smem, lead->stream);
    CUDA_CHECK(cudaEventRecord(lead->ev1, lead->stream));
    for (int i = 0; i < n; ++i) {
      CUDA_CHECK(cudaMemcpyAsync(&hbc[i], es[i]->burst.d_bc, sizeof(BurstCtl), cudaMemcpyDeviceToHost, lead->stream));
      CUDA_CHECK(cudaMemcpyAsync(es[i]->h_block, es[i]->d_block, sizeof(Result), cudaMemcpyDeviceToHost, lead->stream));
    }
    CUDA_CHECK(cudaEventRecord(lead->ev_done, lead->stream));
    for (int i = 1; i < n; ++i) CUDA_CHECK(cudaStreamWaitEvent(es[i]->stream, lead->ev_done, 0));
    CUDA_CHECK(cudaEventRecord(lead->ev_sleep, lead->stream));
    CUDA_CHECK(cudaEventSynchronize(lead->ev_sleep));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, lead->ev0, lead->ev1));
    for (int i = 0; i < n; ++i) es[i]->burst.kernel_seconds += ms * 1e-3 / n;  // (the launch's time, shared out)
  });
  if (rc != PCP_OK) return rc;
  *fused = 1;
  const BurstCtl* hbc = reinterpret_cast<const BurstCtl*>(static_cast<char*>(lead->h_bbatch) +
                                                         (size_t)lead->burst_batch_cap * (sizeof(Params) + sizeof(BurstParams)));
  for (int i = 0; i < n; ++i) {
    int r2 = guarded(es[i], [&] { burst_finish(es[i], hbc[i], &res[i]); });
    if (rc == PCP_OK) rc = r2;
  }
  return rc;
}

// The bit sets of traced nodes [first, first + n) of an IntervalSet engine, re-based into the window
// the caller asks for (as pcp_domains_read_bits does for the current domains).
int pcp_internal_burst_trace_bits(pcp_engine* e, uint64_t first, uint64_t n, const int32_t* lo, const int32_t* hi, int32_t base,
                                  int32_t words, uint32_t* out) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    auto& b = e->burst;
    PCP_REQUIRE(b.open && e->set_mode && b.trace_dom && first + n <= b.trace_cap, "trace range out of bounds");
    if (n == 0) return;
    CUDA_CHECK(cudaSetDevice(e->device));
    const size_t V = e->V, W = (size_t)e->set_W;
    std::vector<uint32_t> raw(n * V * W);
    CUDA_CHECK(cudaMemcpyAsync(raw.data(), b.d_tbits.p + first * V * W, raw.size() * 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    std::memset(out, 0, n * V * (size_t)words * 4);
    for (size_t i = 0; i < n * V; ++i)
      for (long long v = lo[i]; v <= hi[i]; ++v) {
        const unsigned bb = (unsigned)(v - e->set_base);
        if (!((raw[i * W + (bb >> 5)] >> (bb & 31)) & 1u)) continue;
        const long long o = v - base;
        PCP_REQUIRE(o >= 0 && o < (long long)words * 32, "value outside the requested bit window");
        out[i * (size_t)words + (size_t)(o >> 5)] |= 1u << (o & 31);
      }
  });
}

int pcp_internal_burst_trace(pcp_engine* e, uint64_t first, uint64_t n, int32_t* status, int32_t* lo, int32_t* hi) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    auto& b = e->burst;
    PCP_REQUIRE(b.open && first + n <= b.trace_cap, "trace range out of bounds");
    if (n == 0) return;
    CUDA_CHECK(cudaSetDevice(e->device));
    const size_t V = e->V;
    if (status) CUDA_CHECK(cudaMemcpyAsync(status, b.d_tstatus.p + first, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    std::vector<int2> tmp;
    if (lo && hi && b.trace_dom) {
      tmp.resize(n * V);
      CUDA_CHECK(cudaMemcpyAsync(tmp.data(), b.d_tdom.p + first * V, n * V * sizeof(int2), cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < tmp.size(); ++i) { lo[i] = tmp[i].x; hi[i] = tmp[i].y; }
  });
}

int pcp_internal_burst_end(pcp_engine* e) {
  if (!e) return PCP_ERR_INVALID;
  if (!e->burst.open) return PCP_OK;
  // Adopt the device's search state, so that the engine is left where the search stopped
  // (like the space a SearchTreeVisitor hands back): the labels it took, the branching
  // constraints it posted in the tail, the trail.
  return guarded(e, [&] {
    auto& b = e->burst;
    b.open = false;
    CUDA_CHECK(cudaSetDevice(e->device));
    BurstCtl bc;
    CUDA_CHECK(cudaMemcpy(&bc, b.d_bc, sizeof(bc), cudaMemcpyDeviceToHost));
    HostFamily& hb = e->fam[F_BIN];
    const size_t l0 = e->labels.size(), l1 = (size_t)std::max(bc.n_labels, (int)l0);
    if (l1 > l0) {
      std::vector<int2> meta(l1 - l0);
      CUDA_CHECK(cudaMemcpy(meta.data(), b.d_meta.p + l0, meta.size() * sizeof(int2), cudaMemcpyDeviceToHost));
      for (const int2& m : meta) {
        LabelRec r;
        r.n_fam[F_BIN] = (size_t)m.x;
        r.n_fam[F_TER] = e->fam[F_TER].n;
        r.n_fam[F_DJ] = e->fam[F_DJ].n;
        r.n_fam[F_NARY] = e->n_nary;
        r.n_nary_ops = e->h_nary_ops.size();
        r.n_tree_ints = e->h_tree_nodes.size();
        r.n_props = (size_t)(b.props_base + m.x);
        r.trail_len = (unsigned)m.y;
        r.at_fixpoint = true;
        r.dom_version = 0;  // never "the current domains"
        e->labels.push_back(r);
      }
    }
    if ((size_t)bc.bin_n > hb.n) {
      std::vector<int4> tail((size_t)bc.bin_n - hb.n);
      const int4* src = e->shared ? hb.d_tdesc.p + (hb.n - hb.tail_base) : hb.d_desc.p + hb.n;
      CUDA_CHECK(cudaMemcpy(tail.data(), src, tail.size() * sizeof(int4), cudaMemcpyDeviceToHost));
      for (const int4& d : tail) {
        hb.desc.push_back(d);
        e->prop_ref.push_back(make_ref(F_BIN, (unsigned)hb.n++));
      }
      hb.uploaded = hb.active_set = hb.n;
    }
    ++e->dom_version;
    e->mirror_valid = false;
    e->snapshot_valid = false;
    e->at_fixpoint = false;
  });
}

int pcp_internal_interval_set(const pcp_engine* e) { return e && e->set_mode ? 1 : 0; }

// Split-phase pcp_consistency for a host thread that drives several engines (pcp_search_step_many):
// begin = prepare + launch, poll = "has the result arrived?" without blocking, end = collect.
int pcp_internal_consistency_begin(pcp_engine* e) {
  if (!e) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    PCP_REQUIRE_NO_BURST(e);
    CUDA_CHECK(cudaSetDevice(e->device));
    fixpoint_launch(e);
  });
}
int pcp_internal_consistency_poll(pcp_engine* e) {
  if (!e || !e->inflight.active) return 1;
  if (e->inflight.zero_copy) return *(volatile unsigned*)&e->h_result()->seq == e->inflight.P.host_seq ? 1 : 0;
  return cudaStreamQuery(e->stream) == cudaErrorNotReady ? 0 : 1;
}
int pcp_internal_consistency_end(pcp_engine* e, int32_t* status, pcp_stats* stats) {
  if (!e || !status) return PCP_ERR_INVALID;
  return guarded(e, [&] {
    CUDA_CHECK(cudaSetDevice(e->device));
    fixpoint_wait(e, status, stats);
  });
}

int pcp_num_vars(const pcp_engine* e, int32_t* n) {
  if (!e || !n) return PCP_ERR_INVALID;
  *n = (int32_t)e->V;
  return PCP_OK;
}
int pcp_num_props(const pcp_engine* e, int32_t* n) {
  if (!e || !n) return PCP_ERR_INVALID;
  *n = (int32_t)e->num_props();
  return PCP_OK;
}

}  // extern "C"
