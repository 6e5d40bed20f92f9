// pcp_oracle_capi.cpp -- C entry points of the CPU ORACLE (test infrastructure).
//
// Same shapes as include/pcp_b200.h (prefix pcpo_ instead of pcp_) so that the
// parity tests drive the oracle and the device engine with identical inputs.
// See pcp_oracle.hpp for the restatement itself and its reference citations.
#include "pcp_oracle.hpp"

#include <chrono>
#include <cstring>

#include "../include/pcp_b200.h"

using namespace pcpo;

struct pcpo_engine {
  Space space;
  bool flat = false;
  std::vector<std::vector<pcp_operand>> sums;
  std::vector<Label> labels;
  std::string err;
};

namespace {

Var make_view(const pcpo_engine* e, pcp_operand op) {
  if (op.var >= 0) {
    Var id = std::make_unique<Identity>(size_t(op.var));
    if (op.off == 0) return id;
    return std::make_unique<Addition>(std::move(id), op.off);
  }
  if (op.var == PCP_VAR_CONSTANT) return std::make_unique<Constant>(op.off);
  size_t sid = size_t(-2 - op.var);
  PCPO_ASSERT(sid < e->sums.size(), "unknown sum view");
  std::vector<Var> terms;
  for (auto& t : e->sums[sid]) terms.push_back(make_view(e, t));
  Var s = std::make_unique<Sum>(std::move(terms));
  if (op.off == 0) return s;
  return std::make_unique<Addition>(std::move(s), op.off);
}

void check_operand(const pcpo_engine* e, pcp_operand op) {
  if (op.var >= 0) PCPO_ASSERT(size_t(op.var) < e->space.vstore.size(), "operand variable out of range");
  else if (op.var <= -2) PCPO_ASSERT(size_t(-2 - op.var) < e->sums.size(), "unknown sum view");
}

Formula make_prop(const pcpo_engine* e, int kind, const pcp_operand* ops, int n) {
  for (int i = 0; i < n; ++i) check_operand(e, ops[i]);
  if (e->flat) {  // same kind numbering as enum pcp_prop_kind
    std::vector<FOp> f;
    for (int i = 0; i < n; ++i) {
      pcp_operand op = ops[i];
      if (op.var <= -2) {  // single-term sums delegate (term/sum.rs:62-64); wider ones need the view tree
        const auto& terms = e->sums[size_t(-2 - op.var)];
        PCPO_ASSERT(terms.size() == 1, "flat variant: Sum views with more than one term are not supported");
        op = pcp_operand{terms[0].var, terms[0].off + op.off};
      }
      f.push_back(FOp{op.var, op.off});
    }
    static const int arity[10] = {2, 2, 2, 3, 3, 3, -1, 6, 3, -1};
    PCPO_ASSERT(kind >= 0 && kind < 10, "unknown propagator kind");
    PCPO_ASSERT(arity[kind] < 0 ? n >= 1 : n == arity[kind], "arity");
    return make_flat(kind, f.data(), n);
  }
  auto v = [&](int i) { return make_view(e, ops[i]); };
  switch (kind) {
    case PCP_X_LESS_Y: PCPO_ASSERT(n == 2, "arity"); return std::make_unique<XLessY>(v(0), v(1));
    case PCP_X_NEQ_Y: PCPO_ASSERT(n == 2, "arity"); return std::make_unique<XNeqY>(v(0), v(1));
    case PCP_X_EQ_Y: PCPO_ASSERT(n == 2, "arity"); return std::make_unique<XEqY>(v(0), v(1));
    case PCP_X_GREATER_Y_PLUS_Z: PCPO_ASSERT(n == 3, "arity"); return std::make_unique<XGreaterYPlusZ>(v(0), v(1), v(2));
    case PCP_X_LESS_Y_PLUS_Z: PCPO_ASSERT(n == 3, "arity"); return std::make_unique<XLessYPlusZ>(v(0), v(1), v(2));
    case PCP_X_EQ_Y_PLUS_Z: PCPO_ASSERT(n == 3, "arity"); return std::make_unique<XEqYPlusZ>(v(0), v(1), v(2));
    case PCP_X_EQ_Y_MUL_Z: PCPO_ASSERT(n == 3, "arity"); return std::make_unique<XEqYMulZ>(v(0), v(1), v(2));
    case PCP_DISTINCT: {
      PCPO_ASSERT(n >= 1, "arity");
      std::vector<Var> vars;
      for (int i = 0; i < n; ++i) vars.push_back(v(i));
      return std::make_unique<Distinct>(std::move(vars));
    }
    case PCP_ALL_EQUAL: {
      PCPO_ASSERT(n >= 1, "arity");
      std::vector<Var> vars;
      for (int i = 0; i < n; ++i) vars.push_back(v(i));
      return std::make_unique<AllEqual>(std::move(vars));
    }
    case PCP_DISJ2_X_EQ_Y_PLUS_Z: {
      PCPO_ASSERT(n == 6, "arity");
      std::vector<Formula> fs;
      fs.push_back(std::make_unique<XEqYPlusZ>(v(0), v(1), v(2)));
      fs.push_back(std::make_unique<XEqYPlusZ>(v(3), v(4), v(5)));
      return std::make_unique<Disjunction>(std::move(fs));
    }
    default: throw ContractViolation("unknown propagator kind");
  }
}

template <class F>
int guarded(pcpo_engine* e, F&& f) {
  try { f(); return PCP_OK; }
  catch (const ContractViolation& ex) { if (e) e->err = ex.what(); return PCP_ERR_INVALID; }
  catch (const std::bad_alloc&) { if (e) e->err = "out of memory"; return PCP_ERR_NOMEM; }
}

}  // namespace

extern "C" {

int pcpo_engine_create(int variant, pcpo_engine** out) {
  auto* e = new pcpo_engine();
  // 0 faithful | 1 tuned (static CSR, boxed views) | 2 flat (tuned + inline descriptors)
  e->space.cstore.variant = variant == 0 ? Variant::Faithful : Variant::Tuned;
  e->flat = variant == 2;
  *out = e;
  return PCP_OK;
}
void pcpo_engine_destroy(pcpo_engine* e) { delete e; }
const char* pcpo_last_error(const pcpo_engine* e) { return e->err.c_str(); }

int pcpo_vars_alloc(pcpo_engine* e, const int32_t* lo, const int32_t* hi, int32_t n, int32_t* first_idx) {
  return guarded(e, [&] {
    for (int i = 0; i < n; ++i) PCPO_ASSERT(lo[i] <= hi[i], "alloc of an empty domain");
    if (first_idx) *first_idx = int32_t(e->space.vstore.size());
    for (int i = 0; i < n; ++i) e->space.vstore.alloc(Interval(lo[i], hi[i]));
  });
}

int pcpo_sum_alloc(pcpo_engine* e, const pcp_operand* terms, int32_t n, int32_t* sum_id) {
  return guarded(e, [&] {
    PCPO_ASSERT(n >= 1, "At least one variable in sum.");
    for (int i = 0; i < n; ++i) { PCPO_ASSERT(terms[i].var >= -1, "nested sums are not supported"); check_operand(e, terms[i]); }
    e->sums.emplace_back(terms, terms + n);
    if (sum_id) *sum_id = int32_t(e->sums.size() - 1);
  });
}

int pcpo_prop_alloc(pcpo_engine* e, int32_t kind, const pcp_operand* ops, int32_t n_ops, int32_t* idx) {
  return guarded(e, [&] {
    size_t i = e->space.cstore.alloc(make_prop(e, kind, ops, n_ops));
    if (idx) *idx = int32_t(i);
  });
}

int pcpo_props_alloc(pcpo_engine* e, int32_t kind, const pcp_operand* ops, int32_t n_ops, int64_t n_props,
                     int32_t* first_idx) {
  return guarded(e, [&] {
    if (first_idx) *first_idx = int32_t(e->space.cstore.size());
    for (int64_t p = 0; p < n_props; ++p) e->space.cstore.alloc(make_prop(e, kind, ops + p * n_ops, n_ops));
  });
}

int pcpo_consistency(pcpo_engine* e, int32_t* status, pcp_stats* stats) {
  return guarded(e, [&] {
    uint64_t before = e->space.cstore.num_propagations;
    SKleene k = e->space.consistency();
    *status = int32_t(k);
    if (stats) {
      std::memset(stats, 0, sizeof(*stats));
      stats->propagations = e->space.cstore.num_propagations - before;
      uint32_t a = 0;
      for (uint8_t b : e->space.cstore.active) a += b;
      stats->active_props = a;
    }
  });
}

int pcpo_domains_read(pcpo_engine* e, int32_t first, int32_t n, int32_t* lo, int32_t* hi) {
  return guarded(e, [&] {
    PCPO_ASSERT(first >= 0 && n >= 0 && size_t(first) + size_t(n) <= e->space.vstore.size(), "Variable not registered in the store.");
    for (int i = 0; i < n; ++i) { lo[i] = e->space.vstore.memory[first + i].lb; hi[i] = e->space.vstore.memory[first + i].ub; }
  });
}

int pcpo_var_update(pcpo_engine* e, int32_t idx, int32_t lo, int32_t hi, int32_t* ok) {
  return guarded(e, [&] {
    PCPO_ASSERT(idx >= 0, "Variable not registered in the store.");
    bool r = e->space.vstore.update(size_t(idx), Interval(lo, hi));
    e->space.vstore.drain_delta();  // the store starts every consistency() from all-scheduled
    if (ok) *ok = r ? 1 : 0;
  });
}

int pcpo_active_read(pcpo_engine* e, int32_t first, int32_t n, uint8_t* out) {
  return guarded(e, [&] {
    PCPO_ASSERT(first >= 0 && n >= 0 && size_t(first) + size_t(n) <= e->space.cstore.size(), "propagator out of range");
    for (int i = 0; i < n; ++i) out[i] = e->space.cstore.active[first + i];
  });
}

int pcpo_label(pcpo_engine* e, uint64_t* label) {
  return guarded(e, [&] { e->labels.push_back(e->space.label()); *label = e->labels.size() - 1; });
}
int pcpo_restore(pcpo_engine* e, uint64_t label) {
  return guarded(e, [&] {
    PCPO_ASSERT(label < e->labels.size(), "unknown label");
    e->space.restore(e->labels[label]);
    e->labels.resize(label + 1);
  });
}
int pcpo_num_vars(const pcpo_engine* e, int32_t* n) { *n = int32_t(e->space.vstore.size()); return PCP_OK; }
int pcpo_num_props(const pcpo_engine* e, int32_t* n) { *n = int32_t(e->space.cstore.size()); return PCP_OK; }

// The reference's propagator test fixture (propagators/mod.rs:110-129):
// is_subsumed before; propagate; ordered delta; is_subsumed after.
int pcpo_test_propagation(pcpo_engine* e, int32_t prop, int32_t* before, int32_t* propagate_ok,
                          int32_t* after, int32_t* delta /* (var, event) pairs */, int32_t* n_delta) {
  return guarded(e, [&] {
    PCPO_ASSERT(prop >= 0 && size_t(prop) < e->space.cstore.size(), "propagator out of range");
    Propagator& p = *e->space.cstore.propagators[prop];
    VStore& vs = e->space.vstore;
    *before = int32_t(p.is_subsumed(vs));
    bool ok = p.propagate(vs);
    *propagate_ok = ok ? 1 : 0;
    auto d = vs.drain_delta();
    int cap = *n_delta;
    *n_delta = int32_t(d.size());
    for (int i = 0; i < int(d.size()) && i < cap; ++i) { delta[2 * i] = int32_t(d[i].first); delta[2 * i + 1] = int32_t(d[i].second); }
    *after = int32_t(p.is_subsumed(vs));
  });
}

// PropagatorDependencies::dependencies (propagation/ops.rs:27-29).
int pcpo_prop_dependencies(pcpo_engine* e, int32_t prop, int32_t* deps /* (var, event) pairs */, int32_t* n_deps) {
  return guarded(e, [&] {
    PCPO_ASSERT(prop >= 0 && size_t(prop) < e->space.cstore.size(), "propagator out of range");
    auto d = e->space.cstore.propagators[prop]->dependencies();
    int cap = *n_deps;
    *n_deps = int32_t(d.size());
    for (int i = 0; i < int(d.size()) && i < cap; ++i) { deps[2 * i] = int32_t(d[i].first); deps[2 * i + 1] = int32_t(d[i].second); }
  });
}

int pcpo_search_run(pcpo_engine* e, const pcp_search_config* cfg, pcp_search_result* res,
                    int32_t* trace_status, uint64_t* trace_hash, int32_t* trace_lo, int32_t* trace_hi,
                    uint64_t trace_capacity) {
  return guarded(e, [&] {
    Search s;
    s.flat = e->flat;
    s.cfg.node_limit = cfg->node_limit;
    s.cfg.all_solutions = cfg->all_solutions != 0;
    s.cfg.var_sel = cfg->var_sel;
    s.cfg.val_sel = cfg->val_sel;
    s.cfg.distributor = cfg->distributor;
    s.cfg.bb_mode = BBMode(cfg->bb_mode);
    s.cfg.bb_var = size_t(cfg->bb_var);
    uint64_t n = 0;
    size_t V = e->space.vstore.size();
    uint64_t p_warm = e->space.cstore.num_propagations;
    auto t0 = std::chrono::steady_clock::now();
    s.on_node = [&](const Space& sp, int status) {
      if (n + 1 == uint64_t(cfg->warmup_nodes)) { t0 = std::chrono::steady_clock::now(); p_warm = e->space.cstore.num_propagations; }
      if (n < trace_capacity) {
        if (trace_status) trace_status[n] = status;
        if (trace_hash) trace_hash[n] = status == int(False) ? 0 : hash_domains(sp.vstore.memory);
        if (cfg->trace_domains && trace_lo && trace_hi && status != int(False))
          for (size_t i = 0; i < V; ++i) { trace_lo[n * V + i] = sp.vstore.memory[i].lb; trace_hi[n * V + i] = sp.vstore.memory[i].ub; }
      }
      ++n;
    };
    NodeStatus st = s.run(e->space);
    auto t1 = std::chrono::steady_clock::now();
    std::memset(res, 0, sizeof(*res));
    res->status = int32_t(st);
    res->has_bb_value = s.has_bb_value;
    res->bb_value = s.bb_value;
    res->num_nodes = s.stats.num_nodes;
    res->num_solution = s.stats.num_solution;
    res->num_failed_node = s.stats.num_failed_node;
    res->num_prune = s.stats.num_prune;
    res->propagations = e->space.cstore.num_propagations - p_warm;
    res->seconds = std::chrono::duration<double>(t1 - t0).count();
  });
}

// ---- reactor / scheduler handles (for the golden tests of
// reactors/indexed_deps.rs:159-231 and schedulers/relaxed_fifo.rs:78-132) -------
struct pcpo_reactor { IndexedDeps r; };
pcpo_reactor* pcpo_reactor_new(int32_t num_vars, int32_t num_events) { return new pcpo_reactor{IndexedDeps(size_t(num_vars), size_t(num_events))}; }
void pcpo_reactor_free(pcpo_reactor* r) { delete r; }
int pcpo_reactor_subscribe(pcpo_reactor* r, int32_t var, int32_t ev, int32_t prop) {
  return guarded(nullptr, [&] { r->r.subscribe(size_t(var), FDEvent(ev), size_t(prop)); });
}
int pcpo_reactor_unsubscribe(pcpo_reactor* r, int32_t var, int32_t ev, int32_t prop) {
  return guarded(nullptr, [&] { r->r.unsubscribe(size_t(var), FDEvent(ev), size_t(prop)); });
}
int pcpo_reactor_react(pcpo_reactor* r, int32_t var, int32_t ev, int32_t* out, int32_t* n) {
  return guarded(nullptr, [&] {
    auto v = r->r.react(size_t(var), FDEvent(ev));
    int cap = *n;
    *n = int32_t(v.size());
    for (int i = 0; i < int(v.size()) && i < cap; ++i) out[i] = int32_t(v[i]);
  });
}
int pcpo_reactor_is_empty(pcpo_reactor* r) { return r->r.is_empty() ? 1 : 0; }

struct pcpo_fifo { RelaxedFifo f; };
pcpo_fifo* pcpo_fifo_new(int32_t capacity) { return new pcpo_fifo{RelaxedFifo(size_t(capacity))}; }
void pcpo_fifo_free(pcpo_fifo* f) { delete f; }
int pcpo_fifo_schedule(pcpo_fifo* f, int32_t idx) { return guarded(nullptr, [&] { f->f.schedule(size_t(idx)); }); }
int pcpo_fifo_unschedule(pcpo_fifo* f, int32_t idx) { return guarded(nullptr, [&] { f->f.unschedule(size_t(idx)); }); }
int pcpo_fifo_pop(pcpo_fifo* f) { size_t p; return f->f.pop(&p) ? int(p) : -1; }
int pcpo_fifo_is_empty(pcpo_fifo* f) { return f->f.is_empty() ? 1 : 0; }

// FDEvent::new (events/mod.rs:46-70): -1 = None.
int pcpo_event_new(int32_t llo, int32_t lhi, int32_t blo, int32_t bhi, int32_t* ev) {
  return guarded(nullptr, [&] { *ev = event_new(Interval(llo, lhi), Interval(blo, bhi)); });
}

}  // extern "C"
