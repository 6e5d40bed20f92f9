// pcp_oracle_capi.cpp -- C entry points of the CPU ORACLE (test infrastructure).
//
// Same shapes as include/pcp_b200.h (prefix pcpo_ instead of pcp_) so that the
// parity tests drive the oracle and the device engine with identical inputs.
// See pcp_oracle.hpp for the restatement itself and its reference citations.
// The engine exists once per domain instantiation (pcp_oracle_capi_body.hpp): Interval<i32>
// (variants 0 faithful, 1 tuned, 2 flat) and IntervalSet<i32> (the same + 4).
#include "pcp_oracle.hpp"

#include <chrono>
#include <cstring>

#include "../include/pcp_b200.h"

using namespace pcpo;

struct pcpo_engine {
  std::string err;
  virtual ~pcpo_engine() = default;
  virtual void vars_alloc(const int32_t* lo, const int32_t* hi, int32_t n, int32_t* first_idx) = 0;
  virtual void sum_alloc(const pcp_operand* terms, int32_t n, int32_t* sum_id) = 0;
  virtual void props_alloc(int32_t kind, const pcp_operand* ops, int32_t n_ops, int64_t n_props, int32_t* first_idx) = 0;
  virtual void formula_alloc(const int32_t* words, int32_t n_words, int32_t* idx) = 0;
  virtual void consistency(int32_t* status, pcp_stats* stats) = 0;
  virtual void domains_read(int32_t first, int32_t n, int32_t* lo, int32_t* hi) = 0;
  virtual void domains_size_read(int32_t first, int32_t n, uint32_t* size) = 0;
  virtual void domains_read_bits(int32_t first, int32_t n, int32_t base, int32_t words, uint32_t* out) = 0;
  virtual void var_update(int32_t idx, int32_t lo, int32_t hi, int32_t* ok) = 0;
  virtual void active_read(int32_t first, int32_t n, uint8_t* out) = 0;
  virtual void label(uint64_t* l) = 0;
  virtual void restore(uint64_t l) = 0;
  virtual int32_t num_vars() const = 0;
  virtual int32_t num_props() const = 0;
  virtual void test_propagation(int32_t prop, int32_t* before, int32_t* propagate_ok, int32_t* after, int32_t* delta,
                                int32_t* n_delta) = 0;
  virtual void prop_dependencies(int32_t prop, int32_t* deps, int32_t* n_deps) = 0;
  virtual int32_t store_is_subsumed() const = 0;
  virtual void search_run(const pcp_search_config* cfg, pcp_search_result* res, int32_t* trace_status,
                          uint64_t* trace_hash, int32_t* trace_lo, int32_t* trace_hi, uint64_t trace_capacity) = 0;
};

namespace pcpo {
namespace iv {
#include "pcp_oracle_capi_body.hpp"
}
namespace set {
#include "pcp_oracle_capi_body.hpp"
}
}  // namespace pcpo

namespace {
template <class F>
int guarded(pcpo_engine* e, F&& f) {
  try { f(); return PCP_OK; }
  catch (const ContractViolation& ex) { if (e) e->err = ex.what(); return PCP_ERR_INVALID; }
  catch (const std::bad_alloc&) { if (e) e->err = "out of memory"; return PCP_ERR_NOMEM; }
}
}  // namespace

extern "C" {

int pcpo_engine_create(int variant, pcpo_engine** out) {
  // 0 faithful | 1 tuned (static CSR, boxed views) | 2 flat (tuned + inline descriptors); + 4: IntervalSet domains
  const int v = variant & 3;
  if (variant & 4) {
    auto* e = new set::Engine();
    e->space.cstore.variant = v == 0 ? set::Variant::Faithful : set::Variant::Tuned;
    e->flat = v == 2;
    *out = e;
  } else {
    auto* e = new iv::Engine();
    e->space.cstore.variant = v == 0 ? iv::Variant::Faithful : iv::Variant::Tuned;
    e->flat = v == 2;
    *out = e;
  }
  return PCP_OK;
}
void pcpo_engine_destroy(pcpo_engine* e) { delete e; }
const char* pcpo_last_error(const pcpo_engine* e) { return e->err.c_str(); }

int pcpo_vars_alloc(pcpo_engine* e, const int32_t* lo, const int32_t* hi, int32_t n, int32_t* first_idx) {
  return guarded(e, [&] { e->vars_alloc(lo, hi, n, first_idx); });
}
int pcpo_sum_alloc(pcpo_engine* e, const pcp_operand* terms, int32_t n, int32_t* sum_id) {
  return guarded(e, [&] { e->sum_alloc(terms, n, sum_id); });
}
int pcpo_prop_alloc(pcpo_engine* e, int32_t kind, const pcp_operand* ops, int32_t n_ops, int32_t* idx) {
  return guarded(e, [&] { e->props_alloc(kind, ops, n_ops, 1, idx); });
}
int pcpo_props_alloc(pcpo_engine* e, int32_t kind, const pcp_operand* ops, int32_t n_ops, int64_t n_props,
                     int32_t* first_idx) {
  return guarded(e, [&] { e->props_alloc(kind, ops, n_ops, n_props, first_idx); });
}
int pcpo_formula_alloc(pcpo_engine* e, const int32_t* words, int32_t n_words, int32_t* idx) {
  return guarded(e, [&] { e->formula_alloc(words, n_words, idx); });
}
int pcpo_consistency(pcpo_engine* e, int32_t* status, pcp_stats* stats) {
  return guarded(e, [&] { e->consistency(status, stats); });
}
int pcpo_domains_read(pcpo_engine* e, int32_t first, int32_t n, int32_t* lo, int32_t* hi) {
  return guarded(e, [&] { e->domains_read(first, n, lo, hi); });
}
int pcpo_domains_size_read(pcpo_engine* e, int32_t first, int32_t n, uint32_t* size) {
  return guarded(e, [&] { e->domains_size_read(first, n, size); });
}
int pcpo_domains_read_bits(pcpo_engine* e, int32_t first, int32_t n, int32_t base, int32_t words, uint32_t* out) {
  return guarded(e, [&] { e->domains_read_bits(first, n, base, words, out); });
}
int pcpo_var_update(pcpo_engine* e, int32_t idx, int32_t lo, int32_t hi, int32_t* ok) {
  return guarded(e, [&] { e->var_update(idx, lo, hi, ok); });
}
int pcpo_active_read(pcpo_engine* e, int32_t first, int32_t n, uint8_t* out) {
  return guarded(e, [&] { e->active_read(first, n, out); });
}
int pcpo_label(pcpo_engine* e, uint64_t* label) { return guarded(e, [&] { e->label(label); }); }
int pcpo_restore(pcpo_engine* e, uint64_t label) { return guarded(e, [&] { e->restore(label); }); }
int pcpo_num_vars(const pcpo_engine* e, int32_t* n) { *n = e->num_vars(); return PCP_OK; }
int pcpo_num_props(const pcpo_engine* e, int32_t* n) { *n = e->num_props(); return PCP_OK; }
int pcpo_test_propagation(pcpo_engine* e, int32_t prop, int32_t* before, int32_t* propagate_ok,
                          int32_t* after, int32_t* delta /* (var, event) pairs */, int32_t* n_delta) {
  return guarded(e, [&] { e->test_propagation(prop, before, propagate_ok, after, delta, n_delta); });
}
int pcpo_prop_dependencies(pcpo_engine* e, int32_t prop, int32_t* deps /* (var, event) pairs */, int32_t* n_deps) {
  return guarded(e, [&] { e->prop_dependencies(prop, deps, n_deps); });
}
int pcpo_store_is_subsumed(pcpo_engine* e, int32_t* k) { return guarded(e, [&] { *k = e->store_is_subsumed(); }); }
int pcpo_search_run(pcpo_engine* e, const pcp_search_config* cfg, pcp_search_result* res,
                    int32_t* trace_status, uint64_t* trace_hash, int32_t* trace_lo, int32_t* trace_hi,
                    uint64_t trace_capacity) {
  return guarded(e, [&] { e->search_run(cfg, res, trace_status, trace_hash, trace_lo, trace_hi, trace_capacity); });
}

// ---- reactor / scheduler handles (for the golden tests of
// reactors/indexed_deps.rs:159-231 and schedulers/relaxed_fifo.rs:78-132) -------
struct pcpo_reactor { IndexedDeps r; };
pcpo_reactor* pcpo_reactor_new(int32_t num_vars, int32_t num_events) { return new pcpo_reactor{IndexedDeps(size_t(num_vars), size_t(num_events))}; }
void pcpo_reactor_free(pcpo_reactor* r) { delete r; }
int pcpo_reactor_subscribe(pcpo_reactor* r, int32_t var, int32_t ev, int32_t prop) {
  return guarded((pcpo_engine*)nullptr, [&] { r->r.subscribe(size_t(var), FDEvent(ev), size_t(prop)); });
}
int pcpo_reactor_unsubscribe(pcpo_reactor* r, int32_t var, int32_t ev, int32_t prop) {
  return guarded((pcpo_engine*)nullptr, [&] { r->r.unsubscribe(size_t(var), FDEvent(ev), size_t(prop)); });
}
int pcpo_reactor_react(pcpo_reactor* r, int32_t var, int32_t ev, int32_t* out, int32_t* n) {
  return guarded((pcpo_engine*)nullptr, [&] {
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
int pcpo_fifo_schedule(pcpo_fifo* f, int32_t idx) { return guarded((pcpo_engine*)nullptr, [&] { f->f.schedule(size_t(idx)); }); }
int pcpo_fifo_unschedule(pcpo_fifo* f, int32_t idx) { return guarded((pcpo_engine*)nullptr, [&] { f->f.unschedule(size_t(idx)); }); }
int pcpo_fifo_pop(pcpo_fifo* f) { size_t p; return f->f.pop(&p) ? int(p) : -1; }
int pcpo_fifo_is_empty(pcpo_fifo* f) { return f->f.is_empty() ? 1 : 0; }

// FDEvent::new (events/mod.rs:46-70): -1 = None.
int pcpo_event_new(int32_t llo, int32_t lhi, int32_t blo, int32_t bhi, int32_t* ev) {
  return guarded((pcpo_engine*)nullptr, [&] { *ev = iv::event_new(Interval(llo, lhi), Interval(blo, bhi)); });
}

}  // extern "C"
