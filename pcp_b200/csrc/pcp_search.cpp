// pcp_search.cpp -- host search driver over the public C ABI.
//
// The callers of the hot path (reference src/libpcp/search/): OneSolution / AllSolution
// (engine/one_solution.rs:92-105, engine/all_solution.rs:41-50) o StopNode
// (stop_node.rs:54-61) o [BranchAndBound (branch_and_bound.rs:69-94)] o Propagation
// (propagation.rs:42-55) o Brancher(FirstSmallestVar | InputOrder, MiddleVal | MinVal,
// BinarySplit | Enumerate) (branching/*.rs).  It touches the engine exclusively through
// the entry points a Rust host would bind (pcp_restore, pcp_prop_alloc, pcp_consistency,
// pcp_domains_read, pcp_label) -- one fixpoint launch per node, host buffers on both
// sides -- so its timings are end-to-end timings of the boundary.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/pcp_b200.h"
#include "pcp_internal.h"

namespace {

struct Branch {
  uint64_t label;
  int kind;  // 0: x <= v, 1: x > v (binary_split.rs:46-57); 2: x == v, 3: x != v (enumerate.rs)
  int32_t var, val;
};

// FNV-1a over the (lb, ub) words of every maximal run of values of every domain, from a bit window
// (IntervalSet engines; a domain without holes hashes like its (lo, hi) pair below)
uint64_t hash_domain_bits(const int32_t* lo, const int32_t* hi, size_t n, int32_t base, int32_t words, const uint32_t* bits) {
  uint64_t h = 1469598103934665603ull;
  auto feed = [&](uint32_t w) { for (int b = 0; b < 4; ++b) { h ^= (w >> (8 * b)) & 0xffu; h *= 1099511628211ull; } };
  for (size_t i = 0; i < n; ++i) {
    const uint32_t* w = bits + i * (size_t)words;
    auto has = [&](int64_t v) { const int64_t o = v - base; return (w[o >> 5] >> (o & 31)) & 1u; };
    int64_t v = lo[i];
    while (v <= hi[i]) {
      if (!has(v)) { ++v; continue; }
      int64_t u = v;
      while (u + 1 <= hi[i] && has(u + 1)) ++u;
      feed((uint32_t)(int32_t)v);
      feed((uint32_t)(int32_t)u);
      v = u + 1;
    }
  }
  return h;
}

uint64_t hash_domains(const int32_t* lo, const int32_t* hi, size_t n) {  // FNV-1a over (lo,hi) words
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) {
    uint32_t w[2] = {(uint32_t)lo[i], (uint32_t)hi[i]};
    for (int k = 0; k < 2; ++k)
      for (int b = 0; b < 4; ++b) { h ^= (w[k] >> (8 * b)) & 0xffu; h *= 1099511628211ull; }
  }
  return h;
}

#define TRY(expr) do { int _rc = (expr); if (_rc != PCP_OK) return _rc; } while (0)

struct Driver {
  pcp_engine* e;
  const pcp_search_config* cfg;
  pcp_search_result* res;
  int32_t *t_status, *t_lo, *t_hi;
  uint64_t* t_hash;
  uint64_t t_cap;
  std::vector<Branch> queue;  // VectorStack: DFS, left child first
  std::vector<int32_t> lo, hi;
  std::vector<uint32_t> size, bits;  // Cardinality::size() per variable; bit window of the traced domains
  bool set_domains = false;          // IntervalSet engine: sizes and hashes come from the sets
  bool started = false;
  uint64_t nodes_explored = 0;
  int32_t V = 0;
  std::chrono::steady_clock::time_point t_start;

  int apply_alternative(const Branch& b) {
    pcp_operand x{b.var, 0}, c{PCP_VAR_CONSTANT, b.val}, c1{PCP_VAR_CONSTANT, b.val + 1};
    pcp_operand ops[2];
    int kind;
    switch (b.kind) {
      case 0: ops[0] = x; ops[1] = c1; kind = PCP_X_LESS_Y; break;   // x_leq_y(x, v) = x < v + 1 (cmp/mod.rs:54-60)
      case 1: ops[0] = c; ops[1] = x; kind = PCP_X_LESS_Y; break;    // x_greater_y(x, v) = v < x (cmp/mod.rs:40-42)
      case 2: ops[0] = x; ops[1] = c; kind = PCP_X_EQ_Y; break;
      default: ops[0] = x; ops[1] = c; kind = PCP_X_NEQ_Y; break;
    }
    return pcp_prop_alloc(e, kind, ops, 2, nullptr);
  }

  // Propagation::enter + Brancher::enter (+ BranchAndBound, StopNode, Monitor/Statistics), in
  // three steps so that several drivers can put their fixpoints on the GPU together
  // (pcp_search_step_many): enter_pre posts what the node needs, the fixpoint runs
  // (pcp_consistency here, pcp_consistency_batch there), enter_post reads the result and branches.
  // *out: 1 Satisfiable, -1 Unsatisfiable, 0 Unknown (branches pushed), 2 EndOfSearch.
  int enter_pre() {
    if (cfg->bb_mode != 0 && res->has_bb_value) {  // branch_and_bound.rs:76-87
      pcp_operand v{cfg->bb_var, 0}, b{PCP_VAR_CONSTANT, res->bb_value};
      pcp_operand ops[2];
      if (cfg->bb_mode == 1) { ops[0] = v; ops[1] = b; } else { ops[0] = b; ops[1] = v; }
      TRY(pcp_prop_alloc(e, PCP_X_LESS_Y, ops, 2, nullptr));
    }
    if (res->num_nodes == (uint64_t)cfg->warmup_nodes) t_start = std::chrono::steady_clock::now();
    return PCP_OK;
  }
  int enter_child(int* out) {
    TRY(enter_pre());
    int32_t k = 0;
    pcp_stats st;
    TRY(pcp_consistency(e, &k, &st));  // propagation.rs:49
    return enter_post(k, st, out);
  }
  int enter_post(int32_t k, const pcp_stats& st, int* out) {
    if (res->num_nodes >= (uint64_t)cfg->warmup_nodes) {
      res->propagations += st.propagations;
      res->iterations += st.iterations;
      res->kernel_seconds += st.kernel_ms * 1e-3;
    }
    int status = k;
    if (k != PCP_FALSE) TRY(pcp_domains_read(e, 0, V, lo.data(), hi.data()));
    Branch left{}, right{};
    if (k == PCP_UNKNOWN) {
      // first_smallest_var.rs:30-39 / input_order.rs; middle_val.rs:25-27 / min_val.rs
      int32_t var = -1;
      uint32_t best = 0;
      if (set_domains) TRY(pcp_domains_size_read(e, 0, V, size.data()));
      for (int32_t i = 0; i < V; ++i) {
        uint32_t sz = set_domains ? size[i] : (uint32_t)(hi[i] - lo[i]) + 1u;
        if (sz > 1 && (var < 0 || (cfg->var_sel == 0 && sz < best))) { var = i; best = sz; if (cfg->var_sel == 1) break; }
      }
      if (var < 0) return PCP_ERR_INVALID;  // "Cannot select a variable in a space where all variables are assigned."
      int32_t val = cfg->val_sel == 0 ? (lo[var] + hi[var]) / 2 : lo[var];
      uint64_t label = 0;
      TRY(pcp_label(e, &label));  // Branch::distribute (branch.rs:36-49)
      int k0 = cfg->distributor == 0 ? 0 : 2;
      left = Branch{label, k0, var, val};
      right = Branch{label, k0 + 1, var, val};
    }
    uint64_t n = res->num_nodes;
    if (n < t_cap) {
      if (t_status) t_status[n] = k;
      if (t_hash && k != PCP_FALSE && set_domains) {
        int32_t mn = lo[0], mx = hi[0];
        for (int32_t i = 1; i < V; ++i) { mn = std::min(mn, lo[i]); mx = std::max(mx, hi[i]); }
        const int32_t words = (int32_t)(((int64_t)mx - mn + 32) / 32);
        bits.resize((size_t)V * (size_t)words);
        TRY(pcp_domains_read_bits(e, 0, V, mn, words, bits.data()));
        t_hash[n] = hash_domain_bits(lo.data(), hi.data(), (size_t)V, mn, words, bits.data());
      } else if (t_hash) t_hash[n] = k == PCP_FALSE ? 0 : hash_domains(lo.data(), hi.data(), (size_t)V);
      if (cfg->trace_domains && t_lo && t_hi && k != PCP_FALSE) {
        std::memcpy(t_lo + n * (size_t)V, lo.data(), sizeof(int32_t) * (size_t)V);
        std::memcpy(t_hi + n * (size_t)V, hi.data(), sizeof(int32_t) * (size_t)V);
      }
    }
    if (status == PCP_TRUE && cfg->bb_mode != 0) {  // branch_and_bound.rs:89-92
      res->has_bb_value = 1;
      res->bb_value = lo[cfg->bb_var];
    }
    ++nodes_explored;  // stop_node.rs:54-61
    bool stop = cfg->node_limit && nodes_explored >= cfg->node_limit;
    ++res->num_nodes;  // monitor.rs:61-66 sees the status after StopNode
    if (stop) status = 2;
    else if (status == PCP_TRUE) ++res->num_solution;
    else if (status == PCP_FALSE) ++res->num_failed_node;
    if (status == PCP_UNKNOWN) {  // one_solution.rs:46-51: reversed push => left first
      queue.push_back(right);
      queue.push_back(left);
    }
    *out = status;
    return PCP_OK;
  }

  bool stopped = false, exhausted_reported = false;
  std::chrono::steady_clock::time_point t_mark;
  // device-resident bursts (the default search configuration): the node loop below runs on the GPU
  bool burst = false, burst_begun = false;
  pcp_burst_result br_prev{};

  int copy_burst_trace(uint64_t from, uint64_t to) {
    if (to > t_cap) to = t_cap;
    std::vector<int32_t> blo, bhi;
    for (uint64_t a = from; a < to; a += 256) {
      uint64_t n = std::min<uint64_t>(256, to - a);
      int32_t* plo = nullptr;
      int32_t* phi = nullptr;
      if (cfg->trace_domains && t_lo && t_hi) { plo = t_lo + a * (size_t)V; phi = t_hi + a * (size_t)V; }
      else if (t_hash) { blo.resize(n * (size_t)V); bhi.resize(n * (size_t)V); plo = blo.data(); phi = bhi.data(); }
      std::vector<int32_t> st(n);
      TRY(pcp_internal_burst_trace(e, a, n, st.data(), plo, phi));
      const bool sets = pcp_internal_interval_set(e) != 0;
      for (uint64_t i = 0; i < n; ++i) {
        if (t_status) t_status[a + i] = st[i];
        if (!t_hash) continue;
        if (st[i] == PCP_FALSE) { t_hash[a + i] = 0; continue; }
        const int32_t* l = plo + i * (size_t)V;
        const int32_t* h = phi + i * (size_t)V;
        if (!sets) { t_hash[a + i] = hash_domains(l, h, (size_t)V); continue; }
        int32_t mn = l[0], mx = h[0];
        for (int32_t v = 1; v < V; ++v) { mn = std::min(mn, l[v]); mx = std::max(mx, h[v]); }
        const int32_t words = (int32_t)(((int64_t)mx - mn + 32) / 32);
        bits.resize((size_t)V * (size_t)words);
        TRY(pcp_internal_burst_trace_bits(e, a + i, 1, l, h, mn, words, bits.data()));
        t_hash[a + i] = hash_domain_bits(l, h, (size_t)V, mn, words, bits.data());
      }
    }
    return PCP_OK;
  }

  // The device search in slices.  burst_slice_pre: open it on first use and size the next slice
  // (the warm-up nodes end a slice of their own); burst_slice_post: account for what the slice ran;
  // *finished when the call's budget is used up or the search reported a status.
  int burst_slice_pre(uint64_t remaining, uint64_t* budget, bool* timed) {
    if (!burst_begun) {
      TRY(pcp_num_vars(e, &V));
      TRY(pcp_internal_burst_begin(e, cfg->all_solutions, cfg->node_limit, t_cap, cfg->trace_domains));
      burst_begun = true;
    }
    const uint64_t warm = (uint64_t)cfg->warmup_nodes;
    *timed = res->num_nodes >= warm;
    *budget = remaining;
    if (!*timed) *budget = std::min<uint64_t>(*budget, warm - res->num_nodes);
    return PCP_OK;
  }
  int burst_slice_post(const pcp_burst_result& br, double dt, bool timed, bool unlimited, uint64_t* remaining, int* out, bool* finished) {
    const uint64_t ran = br.nodes - br_prev.nodes;
    if (timed) {
      res->seconds += dt;
      res->propagations += br.propagations - br_prev.propagations;
      res->iterations += br.iterations - br_prev.iterations;
      res->kernel_seconds += br.kernel_seconds - br_prev.kernel_seconds;
    }
    if (t_cap && br_prev.nodes < t_cap) TRY(copy_burst_trace(br_prev.nodes, br.nodes));
    res->num_nodes = br.nodes;
    res->num_solution = br.solutions;
    res->num_failed_node = br.failures;
    br_prev = br;
    *finished = false;
    if (br.status != 0) { *out = br.status; if (br.status == 2) stopped = true; *finished = true; return PCP_OK; }
    if (!unlimited) {
      *remaining -= std::min(*remaining, ran);
      if (*remaining == 0) { *out = 0; *finished = true; }
    }
    return PCP_OK;
  }
  int step_burst(uint64_t max_nodes, int* out) {
    uint64_t remaining = max_nodes;
    const bool unlimited = max_nodes == ~0ull;
    while (true) {
      uint64_t budget = 0;
      bool timed = false, finished = false;
      TRY(burst_slice_pre(remaining, &budget, &timed));
      pcp_burst_result br{};
      auto t0 = std::chrono::steady_clock::now();
      TRY(pcp_internal_burst_step(e, budget, &br));
      double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      TRY(burst_slice_post(br, dt, timed, unlimited, &remaining, out, &finished));
      if (finished) return PCP_OK;
    }
  }

  // OneSolution::enter (one_solution.rs:92-105) / AllSolution::enter (all_solution.rs:41-50),
  // resumable: runs at most `max_nodes` further nodes.
  // *out: 0 budget slice used up (search still open), 1 Satisfiable (one-solution mode; call
  // again for the next solution), -1 Unsatisfiable (tree exhausted, one-solution mode),
  // 2 EndOfSearch (StopNode limit, or tree exhausted in all-solutions mode).
  int step(uint64_t max_nodes, int* out) {
    if (burst) {
      int rc = step_burst(max_nodes, out);
      res->status = *out;
      return rc;
    }
    t_mark = std::chrono::steady_clock::now();
    t_start = t_mark;
    int rc = step_inner(max_nodes, out);
    if (res->num_nodes > (uint64_t)cfg->warmup_nodes)
      res->seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    res->status = *out;
    return rc;
  }

  int step_inner(uint64_t max_nodes, int* out) {
    const uint64_t start = res->num_nodes;
    if (V == 0 && !started) {
      TRY(pcp_num_vars(e, &V));
      lo.assign((size_t)V, 0);
      hi.assign((size_t)V, 0);
      size.assign((size_t)V, 0);
      set_domains = pcp_internal_interval_set(e) != 0;
    }
    if (stopped) { *out = 2; return PCP_OK; }
    while (true) {
      if (started && queue.empty()) {  // fully explored (one_solution.rs:66-68)
        if (cfg->all_solutions || exhausted_reported) *out = 2;
        else { exhausted_reported = true; *out = -1; }
        return PCP_OK;
      }
      if (res->num_nodes - start >= max_nodes) { *out = 0; return PCP_OK; }
      int child = 0;
      if (!started) {
        started = true;
      } else {
        Branch b = queue.back();
        queue.pop_back();
        TRY(pcp_restore(e, b.label));  // Branch::commit (branch.rs:51-55)
        TRY(apply_alternative(b));
      }
      TRY(enter_child(&child));
      if (child == 2) { stopped = true; *out = 2; return PCP_OK; }
      if (child == 1 && !cfg->all_solutions) { *out = 1; return PCP_OK; }
    }
  }
};

}  // namespace

// A worker thread that belongs to one search: pcp_search_step_many hands it slices of nodes, so
// that several searches (one engine, one host thread each -- the threading contract of the
// header) advance at the same time without a thread being created per call.
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  bool has_job = false, done = false, quit = false;
  uint64_t max_nodes = 0;
  int rc = PCP_OK;
  std::vector<pcp_search*> group;  // the searches this thread advances (host-driven: pipelined over them)
};

struct pcp_search {
  pcp_search_config cfg;
  pcp_search_result res;
  Driver d;
  Worker* worker = nullptr;
};

extern "C" {

int pcp_search_open(pcp_engine* e, const pcp_search_config* cfg, int32_t* trace_status, uint64_t* trace_hash,
                    int32_t* trace_lo, int32_t* trace_hi, uint64_t trace_capacity, pcp_search** out) {
  if (!e || !cfg || !out) return PCP_ERR_INVALID;
  pcp_search* s = new pcp_search{*cfg, {}, {}};
  std::memset(&s->res, 0, sizeof(s->res));
  s->d = Driver{e, &s->cfg, &s->res, trace_status, trace_lo, trace_hi, trace_hash, trace_capacity};
  s->d.burst = pcp_internal_burst_supported(e, cfg, trace_capacity) != 0;
  *out = s;
  return PCP_OK;
}

int pcp_search_step(pcp_search* s, uint64_t max_nodes, pcp_search_result* res) {
  if (!s) return PCP_ERR_INVALID;
  int status = 0;
  int rc = s->d.step(max_nodes ? max_nodes : ~0ull, &status);
  if (res) *res = s->res;
  return rc;
}

// One host thread, several host-driven searches: every search always has its next fixpoint in
// flight; the thread goes round, collects what has finished (pcp_internal_consistency_poll reads
// the result's sequence word in the mapped mirror), does that node's host work -- read the domains,
// branch, label, restore, post -- and launches the next one.  The searches never wait for each
// other (no rounds), and a host with fewer cores than contexts still keeps every context busy.
static int run_group_pipelined(const std::vector<pcp_search*>& g, uint64_t max_nodes) {
  const uint64_t budget = max_nodes ? max_nodes : ~0ull;
  const size_t n = g.size();
  std::vector<uint64_t> start(n);
  std::vector<char> done(n, 0), inflight(n, 0);
  std::vector<std::chrono::steady_clock::time_point> t0(n);
  for (size_t i = 0; i < n; ++i) {
    Driver& d = g[i]->d;
    start[i] = d.res->num_nodes;
    if (d.V == 0 && !d.started) {
      TRY(pcp_num_vars(d.e, &d.V));
      d.lo.assign((size_t)d.V, 0);
      d.hi.assign((size_t)d.V, 0);
      d.size.assign((size_t)d.V, 0);
      d.set_domains = pcp_internal_interval_set(d.e) != 0;
    }
    if (d.stopped) { d.res->status = 2; done[i] = 1; }
    t0[i] = std::chrono::steady_clock::now();
  }
  size_t open_n = 0;
  for (size_t i = 0; i < n; ++i) open_n += done[i] ? 0 : 1;
  // one visit of search i: collect its finished fixpoint (if any), then launch its next node
  auto visit = [&](size_t i) -> int {
    Driver& d = g[i]->d;
    if (inflight[i]) {
      if (!pcp_internal_consistency_poll(d.e)) return PCP_OK;
      int32_t k = 0;
      pcp_stats st;
      inflight[i] = 0;
      TRY(pcp_internal_consistency_end(d.e, &k, &st));
      int child = 0;
      TRY(d.enter_post(k, st, &child));
      if (child == 2) { d.stopped = true; d.res->status = 2; done[i] = 1; }
      else if (child == 1 && !d.cfg->all_solutions) { d.res->status = 1; done[i] = 1; }
      if (done[i]) { --open_n; return PCP_OK; }
    }
    if (d.started && d.queue.empty()) {  // fully explored (one_solution.rs:66-68)
      if (d.cfg->all_solutions || d.exhausted_reported) d.res->status = 2;
      else { d.exhausted_reported = true; d.res->status = -1; }
      done[i] = 1; --open_n;
      return PCP_OK;
    }
    if (d.res->num_nodes - start[i] >= budget) { d.res->status = 0; done[i] = 1; --open_n; return PCP_OK; }
    if (!d.started) {
      d.started = true;
    } else {
      Branch b = d.queue.back();
      d.queue.pop_back();
      TRY(pcp_restore(d.e, b.label));  // Branch::commit (branch.rs:51-55)
      TRY(d.apply_alternative(b));
    }
    TRY(d.enter_pre());
    TRY(pcp_internal_consistency_begin(d.e));
    inflight[i] = 1;
    return PCP_OK;
  };
  int rc = PCP_OK;
  while (open_n > 0 && rc == PCP_OK)
    for (size_t i = 0; i < n && rc == PCP_OK; ++i)
      if (!done[i]) rc = visit(i);
  if (rc != PCP_OK) {  // leave no fixpoint in flight behind an error
    for (size_t i = 0; i < n; ++i)
      if (inflight[i]) { int32_t k; pcp_internal_consistency_end(g[i]->d.e, &k, nullptr); }
    return rc;
  }
  const auto t1 = std::chrono::steady_clock::now();
  for (size_t i = 0; i < n; ++i) {
    Driver& d = g[i]->d;
    if (d.res->num_nodes > (uint64_t)d.cfg->warmup_nodes) d.res->seconds += std::chrono::duration<double>(t1 - t0[i]).count();
  }
  return PCP_OK;
}

// Several searches -- independent subtrees of one model, one engine each -- advanced together:
// device-resident searches run on one host thread each (their launches last for whole slices of
// nodes); host-driven searches move in lockstep, one node per search and round, the fixpoints of
// a round launched together through pcp_consistency_batch.
int pcp_search_step_many(pcp_search* const* ss, int32_t n, uint64_t max_nodes, pcp_search_result* res) {
  if (!ss || n < 0) return PCP_ERR_INVALID;
  for (int i = 0; i < n; ++i) if (!ss[i]) return PCP_ERR_INVALID;
  if (n == 0) return PCP_OK;
  const uint64_t budget = max_nodes ? max_nodes : ~0ull;
  bool all_burst = true, any_burst = false;
  for (int i = 0; i < n; ++i) { all_burst &= ss[i]->d.burst; any_burst |= ss[i]->d.burst; }
  if (any_burst && !all_burst) return PCP_ERR_INVALID;
  // Host-driven searches can move in lockstep on the calling thread (PCP_SEARCH_LOCKSTEP=1: one
  // pcp_consistency_batch per round) or -- the default, like the device-resident ones -- each on
  // its own worker thread, so that a search that needs a worklist iteration does not hold up the
  // round and the host work of one node (post, launch, read back, branch) overlaps the others'.
  static const bool lockstep = [] { const char* v = std::getenv("PCP_SEARCH_LOCKSTEP"); return v && v[0] == '1'; }();
  if (all_burst && n > 1) {
    // Device-resident searches whose engines can share a launch (forks of one model: same kernel, same
    // geometry) advance slice by slice in ONE launch -- a group of CTAs per search, no host thread per
    // search, every search free-running through its budget inside the launch.
    const bool unlimited = budget == ~0ull;
    std::vector<uint64_t> remaining((size_t)n, budget), budgets;
    std::vector<char> fin((size_t)n, 0), timed;
    std::vector<int> who, outs((size_t)n, 0);
    std::vector<pcp_engine*> es;
    std::vector<pcp_burst_result> brs;
    bool first = true, fell_back = false;
    while (true) {
      who.clear(); es.clear(); budgets.clear(); timed.clear();
      for (int i = 0; i < n; ++i) {
        if (fin[(size_t)i]) continue;
        Driver& d = ss[i]->d;
        if (d.stopped) { outs[(size_t)i] = 2; fin[(size_t)i] = 1; continue; }
        uint64_t b = 0;
        bool t = false;
        TRY(d.burst_slice_pre(remaining[(size_t)i], &b, &t));
        who.push_back(i); es.push_back(d.e); budgets.push_back(b); timed.push_back(t ? 1 : 0);
      }
      if (who.empty()) break;
      if (who.size() == 1 && !first) {  // a straggler: its own launch with the whole grid share it has
        Driver& d = ss[who[0]]->d;
        int o = 0;
        TRY(d.step_burst(remaining[(size_t)who[0]], &o));
        outs[(size_t)who[0]] = o;
        fin[(size_t)who[0]] = 1;
        continue;
      }
      brs.assign(who.size(), pcp_burst_result{});
      int32_t fused = 0;
      auto t0 = std::chrono::steady_clock::now();
      TRY(pcp_internal_burst_step_many(es.data(), (int32_t)es.size(), budgets.data(), brs.data(), &fused));
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (!fused) {
        if (first) { fell_back = true; break; }  // never could: one thread and one launch per search below
        // they no longer can (a search's tail grew past a launch-geometry step): finish this call one by one
        for (size_t k = 0; k < who.size(); ++k) {
          int o = 0;
          TRY(ss[who[k]]->d.step_burst(remaining[(size_t)who[k]], &o));
          outs[(size_t)who[k]] = o;
          fin[(size_t)who[k]] = 1;
        }
        continue;
      }
      first = false;
      for (size_t k = 0; k < who.size(); ++k) {
        Driver& d = ss[who[k]]->d;
        bool f = false;
        int o = 0;
        TRY(d.burst_slice_post(brs[k], dt, timed[k] != 0, unlimited, &remaining[(size_t)who[k]], &o, &f));
        if (f) { outs[(size_t)who[k]] = o; fin[(size_t)who[k]] = 1; }
      }
    }
    if (!fell_back) {
      for (int i = 0; i < n; ++i) {
        ss[i]->d.res->status = outs[(size_t)i];
        ss[i]->res.status = outs[(size_t)i];
        if (res) res[i] = ss[i]->res;
      }
      return PCP_OK;
    }
  }
  if (all_burst || (!lockstep && n > 1)) {
    // device-resident searches: one (sleeping) thread each.  Host-driven searches: T threads, each
    // pipelining its share of the searches; T = PCP_SEARCH_THREADS, else what the machine has.
    int T = n;
    if (!all_burst) {
      const char* v = std::getenv("PCP_SEARCH_THREADS");
      const int hw = (int)std::thread::hardware_concurrency();
      T = v ? std::atoi(v) : (hw > 1 ? hw - 1 : 1);
      T = std::max(1, std::min(T, n));
    }
    // searches are dealt to the leaders' groups round-robin; the leaders are ss[0 .. T)
    std::vector<std::vector<pcp_search*>> groups((size_t)T);
    for (int i = 0; i < n; ++i) groups[(size_t)(i % T)].push_back(ss[i]);
    for (int t = 0; t < T; ++t) {
      pcp_search* s = ss[t];
      if (!s->worker) {
        s->worker = new Worker();
        Worker* w = s->worker;
        w->th = std::thread([s, w] {
          std::unique_lock<std::mutex> lk(w->mu);
          while (true) {
            w->cv.wait(lk, [w] { return w->has_job || w->quit; });
            if (w->quit) return;
            w->has_job = false;
            const uint64_t budget_w = w->max_nodes;
            lk.unlock();
            const int rc = (w->group.size() > 1 || !s->d.burst) ? run_group_pipelined(w->group, budget_w) : pcp_search_step(s, budget_w, nullptr);
            lk.lock();
            w->rc = rc;
            w->done = true;
            w->cv.notify_all();
          }
        });
      }
      Worker* w = s->worker;
      std::lock_guard<std::mutex> lk(w->mu);
      w->group = groups[(size_t)t];
      w->max_nodes = max_nodes;
      w->done = false;
      w->has_job = true;
      w->cv.notify_all();
    }
    int rc = PCP_OK;
    for (int t = 0; t < T; ++t) {
      Worker* w = ss[t]->worker;
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [w] { return w->done; });
      if (w->rc != PCP_OK && rc == PCP_OK) rc = w->rc;
    }
    for (int i = 0; i < n; ++i) {
      ss[i]->res.status = ss[i]->d.res->status;
      if (res) res[i] = ss[i]->res;
    }
    return rc;
  }
  std::vector<uint64_t> start((size_t)n);
  std::vector<char> done((size_t)n, 0);
  for (int i = 0; i < n; ++i) {
    Driver& d = ss[i]->d;
    start[(size_t)i] = d.res->num_nodes;
    if (d.V == 0 && !d.started) {
      TRY(pcp_num_vars(d.e, &d.V));
      d.lo.assign((size_t)d.V, 0);
      d.hi.assign((size_t)d.V, 0);
      d.size.assign((size_t)d.V, 0);
      d.set_domains = pcp_internal_interval_set(d.e) != 0;
    }
    if (d.stopped) { d.res->status = 2; done[(size_t)i] = 1; }
  }
  std::vector<pcp_engine*> engines;
  std::vector<int> who;
  std::vector<int32_t> ks;
  std::vector<pcp_stats> sts;
  while (true) {
    engines.clear();
    who.clear();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i) {
      if (done[(size_t)i]) continue;
      Driver& d = ss[i]->d;
      if (d.started && d.queue.empty()) {  // fully explored (one_solution.rs:66-68)
        if (d.cfg->all_solutions || d.exhausted_reported) d.res->status = 2;
        else { d.exhausted_reported = true; d.res->status = -1; }
        done[(size_t)i] = 1;
        continue;
      }
      if (d.res->num_nodes - start[(size_t)i] >= budget) { d.res->status = 0; done[(size_t)i] = 1; continue; }
      if (!d.started) {
        d.started = true;
      } else {
        Branch b = d.queue.back();
        d.queue.pop_back();
        TRY(pcp_restore(d.e, b.label));  // Branch::commit (branch.rs:51-55)
        TRY(d.apply_alternative(b));
      }
      TRY(d.enter_pre());
      engines.push_back(d.e);
      who.push_back(i);
    }
    if (engines.empty()) break;
    ks.assign(engines.size(), 0);
    sts.assign(engines.size(), pcp_stats{});
    TRY(pcp_consistency_batch(engines.data(), (int32_t)engines.size(), ks.data(), sts.data()));
    for (size_t j = 0; j < who.size(); ++j) {
      Driver& d = ss[who[j]]->d;
      int child = 0;
      TRY(d.enter_post(ks[j], sts[j], &child));
      if (child == 2) { d.stopped = true; d.res->status = 2; done[(size_t)who[j]] = 1; }
      else if (child == 1 && !d.cfg->all_solutions) { d.res->status = 1; done[(size_t)who[j]] = 1; }
    }
    // the wall clock of a round belongs to every search that took part in it
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int i : who) {
      Driver& d = ss[i]->d;
      if (d.res->num_nodes > (uint64_t)d.cfg->warmup_nodes) d.res->seconds += dt;
    }
  }
  if (res) for (int i = 0; i < n; ++i) res[i] = ss[i]->res;
  return PCP_OK;
}

int pcp_search_set_incumbent(pcp_search* s, int32_t value) {
  if (!s || s->cfg.bb_mode == 0) return PCP_ERR_INVALID;
  pcp_search_result& r = s->res;
  const bool better = !r.has_bb_value || (s->cfg.bb_mode == 1 ? value < r.bb_value : value > r.bb_value);
  if (better) { r.has_bb_value = 1; r.bb_value = value; }
  return PCP_OK;
}

void pcp_search_close(pcp_search* s) {
  if (!s) return;
  if (s->worker) {
    { std::lock_guard<std::mutex> lk(s->worker->mu); s->worker->quit = true; s->worker->cv.notify_all(); }
    s->worker->th.join();
    delete s->worker;
  }
  if (s->d.burst_begun) pcp_internal_burst_end(s->d.e);  // the engine goes back to the search root
  delete s;
}

int pcp_search_run(pcp_engine* e, const pcp_search_config* cfg, pcp_search_result* res, int32_t* trace_status,
                   uint64_t* trace_hash, int32_t* trace_lo, int32_t* trace_hi, uint64_t trace_capacity) {
  if (!e || !cfg || !res) return PCP_ERR_INVALID;
  pcp_search* s = nullptr;
  int rc = pcp_search_open(e, cfg, trace_status, trace_hash, trace_lo, trace_hi, trace_capacity, &s);
  if (rc != PCP_OK) return rc;
  // AllSolution keeps calling its child until EndOfSearch; OneSolution returns at the first solution
  rc = pcp_search_step(s, 0, res);
  pcp_search_close(s);
  return rc;
}

}  // extern "C"
