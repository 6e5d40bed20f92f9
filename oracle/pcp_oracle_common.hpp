// pcp_oracle.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A C++17 restatement of the libpcp 0.7.0 propagation fixpoint for the
// `Interval<i32>` instantiation (VStoreFD), written from the reference sources
// under /root/reference (file:line cited at every function).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// build, link or execute anything in this directory; the product (pcp_b200/)
// never does.
//
// Parity pinning: the reference cannot be compiled here (no Rust toolchain), and
// the domain arithmetic lives in the un-vendored crate `intervallum ^1.2.0`
// (Cargo.toml:22-30).  The oracle is pinned by replaying every known-answer
// vector that libpcp's own unit tests hold for this path (tests/golden/*.json,
// transcribed with file:line) -- per-propagator results, delta event lists,
// entailment before/after, reactor/scheduler orderings, search goldens.
// Residual "parity unpinned" items: i32 overflow behaviour and whole-loop
// results at scale (pinned only transitively), see DESIGN.md.
#pragma once  // (this file: what both domain instantiations share)
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pcpo {

struct ContractViolation : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define PCPO_ASSERT(cond, msg) \
  do { if (!(cond)) throw ::pcpo::ContractViolation(msg); } while (0)

// ---------------------------------------------------------------------------
// trilean::SKleene (third-party `trilean ^1.0.1`; strong Kleene logic).
// Pinned by x_eq_y_plus_z.rs:126-140, x_neq_y.rs:124-132.
// ---------------------------------------------------------------------------
enum SKleene : int { False = -1, Unknown = 0, True = 1 };
inline SKleene k_and(SKleene a, SKleene b) { return a < b ? a : b; }
inline SKleene k_or(SKleene a, SKleene b) { return a > b ? a : b; }
inline SKleene k_not(SKleene a) { return static_cast<SKleene>(-static_cast<int>(a)); }

// ---------------------------------------------------------------------------
// intervallum `Interval<i32>`.  empty <=> lb > ub.
// ---------------------------------------------------------------------------
struct Interval {
  int32_t lb, ub;
  Interval() : lb(1), ub(0) {}
  Interval(int32_t l, int32_t u) : lb(l), ub(u) {}
  static Interval empty() { return Interval(1, 0); }
  static Interval singleton(int32_t v) { return Interval(v, v); }
  int32_t lower() const { return lb; }
  int32_t upper() const { return ub; }
  bool is_empty() const { return lb > ub; }
  bool is_singleton() const { return lb == ub; }
  // size as u32, 0 if empty (variable/store.rs:159 compares sizes).
  uint32_t size() const { return is_empty() ? 0u : static_cast<uint32_t>(ub - lb) + 1u; }
  Interval shrink_left(int32_t v) const { return Interval(std::max(lb, v), ub); }
  Interval shrink_right(int32_t v) const { return Interval(lb, std::min(ub, v)); }
  Interval strict_shrink_left(int32_t v) const { return shrink_left(v + 1); }
  Interval strict_shrink_right(int32_t v) const { return shrink_right(v - 1); }
  // Interval \ {v}: only a value sitting on a bound is removed (x_neq_y.rs:128-131).
  Interval difference(int32_t v) const {
    if (is_empty()) return *this;
    if (v == lb) return Interval(lb + 1, ub);
    if (v == ub) return Interval(lb, ub - 1);
    return *this;
  }
  Interval intersection(const Interval& o) const {
    if (is_empty() || o.is_empty()) return empty();
    return Interval(std::max(lb, o.lb), std::min(ub, o.ub));
  }
  bool is_subset(const Interval& o) const {
    if (is_empty()) return true;
    if (o.is_empty()) return false;
    return lb >= o.lb && ub <= o.ub;
  }
  bool is_disjoint(const Interval& o) const {
    return is_empty() || o.is_empty() || lb > o.ub || o.lb > ub;
  }
  bool overlap(const Interval& o) const { return !is_disjoint(o); }
  bool contains(int32_t v) const { return !is_empty() && lb <= v && v <= ub; }
  Interval plus(int32_t c) const { return is_empty() ? empty() : Interval(lb + c, ub + c); }
  Interval minus(int32_t c) const { return is_empty() ? empty() : Interval(lb - c, ub - c); }
  Interval add(const Interval& o) const {
    if (is_empty() || o.is_empty()) return empty();
    return Interval(lb + o.lb, ub + o.ub);
  }
  Interval sub(const Interval& o) const {
    if (is_empty() || o.is_empty()) return empty();
    return Interval(lb - o.ub, ub - o.lb);
  }
  Interval mul(const Interval& o) const {
    if (is_empty() || o.is_empty()) return empty();
    int64_t p[4] = {int64_t(lb) * o.lb, int64_t(lb) * o.ub, int64_t(ub) * o.lb, int64_t(ub) * o.ub};
    return Interval(int32_t(*std::min_element(p, p + 4)), int32_t(*std::max_element(p, p + 4)));
  }
  bool operator==(const Interval& o) const {
    return (is_empty() && o.is_empty()) || (lb == o.lb && ub == o.ub);
  }
  bool operator!=(const Interval& o) const { return !(*this == o); }
  template <class F> void for_each_run(F f) const { if (!is_empty()) f(lb, ub); }   // maximal runs of values
};

// ---------------------------------------------------------------------------
// intervallum `IntervalSet<i32>` (the domain of VStoreSet / FDSpace: variable/mod.rs:38,
// search/mod.rs:41-43, example/src/nqueens.rs:34): a finite set of integers held as a sorted
// vector of disjoint, non-adjacent, non-empty intervals.  The crate is not vendored
// (Cargo.toml:22-30); restated here are the *set* semantics its documentation states for the
// operations libpcp calls (SURVEY 8c lists the call sites): difference(&v) removes one value
// (interior values included -- this is what makes XNeqY stronger than on Interval and raises
// Inner events, events/mod.rs:62-63), shrink_* drops everything below / above a value,
// intersection / is_subset / is_disjoint are the set operations, size() is the cardinality,
// `d + c` shifts, and lower()/upper() are the extreme elements.
// Pinned by libpcp's own tests that run on IntervalSet: binary_split.rs:75-134 (children of
// BinarySplit), first_smallest_var.rs:52-78 (sizes), and the FDSpace search goldens
// (all_solution.rs:67-98 counts, one_solution.rs:120-140 statuses, stop_node.rs:82-104,
// branch_and_bound.rs:112-138) -- the counts also pin that entailment of XNeqY is exact
// enough never to report a solution early.  Anything else about IntervalSet (the order in
// which equal results are represented, `+`/`-`/`*` between two sets -- used only by Sum views
// and XEqYMulZ, which the device refuses on set domains) is "parity unpinned".
// ---------------------------------------------------------------------------
struct IntervalSet {
  std::vector<Interval> iv;
  IntervalSet() = default;
  IntervalSet(int32_t l, int32_t u) { if (l <= u) iv.emplace_back(l, u); }
  static IntervalSet empty() { return IntervalSet(); }
  static IntervalSet singleton(int32_t v) { return IntervalSet(v, v); }
  int32_t lower() const { return iv.empty() ? 1 : iv.front().lb; }   // (empty reads like Interval::empty())
  int32_t upper() const { return iv.empty() ? 0 : iv.back().ub; }
  bool is_empty() const { return iv.empty(); }
  bool is_singleton() const { return iv.size() == 1 && iv[0].lb == iv[0].ub; }
  uint32_t size() const { uint32_t s = 0; for (auto& i : iv) s += i.size(); return s; }
  void push(int32_t l, int32_t u) {  // append in ascending order, joining adjacent runs
    if (l > u) return;
    if (!iv.empty() && int64_t(iv.back().ub) + 1 >= l) iv.back().ub = std::max(iv.back().ub, u);
    else iv.emplace_back(l, u);
  }
  IntervalSet shrink_left(int32_t v) const {
    IntervalSet r;
    for (auto& i : iv) if (i.ub >= v) r.push(std::max(i.lb, v), i.ub);
    return r;
  }
  IntervalSet shrink_right(int32_t v) const {
    IntervalSet r;
    for (auto& i : iv) if (i.lb <= v) r.push(i.lb, std::min(i.ub, v));
    return r;
  }
  IntervalSet strict_shrink_left(int32_t v) const { return shrink_left(v + 1); }
  IntervalSet strict_shrink_right(int32_t v) const { return shrink_right(v - 1); }
  IntervalSet difference(int32_t v) const {  // set \ {v}
    IntervalSet r;
    for (auto& i : iv) {
      if (v < i.lb || v > i.ub) { r.iv.push_back(i); continue; }
      if (i.lb <= v - 1) r.iv.emplace_back(i.lb, v - 1);
      if (v + 1 <= i.ub) r.iv.emplace_back(v + 1, i.ub);
    }
    return r;
  }
  IntervalSet intersection(const IntervalSet& o) const {
    IntervalSet r;
    size_t a = 0, b = 0;
    while (a < iv.size() && b < o.iv.size()) {
      int32_t l = std::max(iv[a].lb, o.iv[b].lb), u = std::min(iv[a].ub, o.iv[b].ub);
      if (l <= u) r.push(l, u);
      if (iv[a].ub < o.iv[b].ub) ++a; else ++b;
    }
    return r;
  }
  bool is_subset(const IntervalSet& o) const {
    size_t b = 0;
    for (auto& i : iv) {
      while (b < o.iv.size() && o.iv[b].ub < i.lb) ++b;
      if (b == o.iv.size() || o.iv[b].lb > i.lb || o.iv[b].ub < i.ub) return false;
    }
    return true;
  }
  bool is_disjoint(const IntervalSet& o) const {
    size_t a = 0, b = 0;
    while (a < iv.size() && b < o.iv.size()) {
      if (std::max(iv[a].lb, o.iv[b].lb) <= std::min(iv[a].ub, o.iv[b].ub)) return false;
      if (iv[a].ub < o.iv[b].ub) ++a; else ++b;
    }
    return true;
  }
  bool overlap(const IntervalSet& o) const { return !is_disjoint(o); }
  bool contains(int32_t v) const { for (auto& i : iv) if (i.lb <= v && v <= i.ub) return true; return false; }
  IntervalSet plus(int32_t c) const { IntervalSet r = *this; for (auto& i : r.iv) { i.lb += c; i.ub += c; } return r; }
  IntervalSet minus(int32_t c) const { return plus(-c); }
  // set (op) set = union over the pairs of intervals of interval (op) interval  [parity unpinned]
  template <class F>
  IntervalSet pairwise(const IntervalSet& o, F f) const {
    std::vector<Interval> all;
    for (auto& i : iv) for (auto& j : o.iv) all.push_back(f(i, j));
    std::sort(all.begin(), all.end(), [](const Interval& x, const Interval& y) { return x.lb < y.lb; });
    IntervalSet r;
    for (auto& i : all) r.push(i.lb, i.ub);
    return r;
  }
  IntervalSet add(const IntervalSet& o) const { return pairwise(o, [](const Interval& i, const Interval& j) { return i.add(j); }); }
  IntervalSet sub(const IntervalSet& o) const { return pairwise(o, [](const Interval& i, const Interval& j) { return i.sub(j); }); }
  IntervalSet mul(const IntervalSet& o) const { return pairwise(o, [](const Interval& i, const Interval& j) { return i.mul(j); }); }
  bool operator==(const IntervalSet& o) const {
    if (iv.size() != o.iv.size()) return false;
    for (size_t k = 0; k < iv.size(); ++k) if (iv[k].lb != o.iv[k].lb || iv[k].ub != o.iv[k].ub) return false;
    return true;
  }
  bool operator!=(const IntervalSet& o) const { return !(*this == o); }
  template <class F> void for_each_run(F f) const { for (auto& i : iv) f(i.lb, i.ub); }
};
// ---------------------------------------------------------------------------
// propagation/events/mod.rs:24-70 -- FDEvent, merge = min, MonotonicEvent::new.
// ---------------------------------------------------------------------------
enum FDEvent : int { Assignment = 0, Bound = 1, Inner = 2 };
constexpr int kNumEvents = 3;                                   // events/mod.rs:41-43
inline FDEvent merge(FDEvent e, FDEvent f) { return e < f ? e : f; }  // events/mod.rs:30-34
using Dep = std::pair<size_t, FDEvent>;

// ---------------------------------------------------------------------------
// propagation/reactors/indexed_deps.rs:23-113
// ---------------------------------------------------------------------------
struct IndexedDeps {
  size_t num_events = 0, num_subscriptions = 0;
  std::vector<std::vector<size_t>> deps;
  bool check_duplicates = true;  // the release-mode assert! at indexed_deps.rs:69-77
  IndexedDeps() = default;
  IndexedDeps(size_t num_vars, size_t nev) : num_events(nev), deps(num_vars * nev) {}
  size_t num_vars() const { return num_events ? deps.size() / num_events : 0; }
  void subscribe(size_t var, FDEvent ev, size_t prop) {
    if (check_duplicates) {
      for (size_t e = 0; e < num_events && var * num_events + e < deps.size(); ++e)
        for (size_t p : deps[var * num_events + e])
          PCPO_ASSERT(p != prop, "propagator already subscribed to this variable");
    }
    PCPO_ASSERT(var < num_vars(), "Reactor IndexedDeps: subscription out of range");
    ++num_subscriptions;
    deps[num_events * var + ev].push_back(prop);
  }
  void unsubscribe(size_t var, FDEvent ev, size_t prop) {
    PCPO_ASSERT(var < num_vars(), "Reactor IndexedDeps: unsubscription out of range");
    auto& props = deps[num_events * var + ev];
    auto it = std::find(props.begin(), props.end(), prop);
    PCPO_ASSERT(it != props.end(), "cannot unsubscribe propagator not registered.");
    --num_subscriptions;
    *it = props.back();  // Vec::swap_remove
    props.pop_back();
  }
  std::vector<size_t> react(size_t var, FDEvent ev) const {
    PCPO_ASSERT(var < num_vars(), "Reactor IndexedDeps: react out of range");
    std::vector<size_t> out;
    for (size_t e = ev; e < num_events; ++e)
      for (size_t p : deps[num_events * var + e]) out.push_back(p);
    return out;
  }
  bool is_empty() const { return num_subscriptions == 0; }
};

// ---------------------------------------------------------------------------
// propagation/schedulers/relaxed_fifo.rs:27-71
// ---------------------------------------------------------------------------
struct RelaxedFifo {
  std::vector<uint8_t> inside_queue;
  std::deque<size_t> queue;
  size_t capacity = 0;
  RelaxedFifo() = default;
  explicit RelaxedFifo(size_t cap) : inside_queue(cap, 0), capacity(cap) {}
  void schedule(size_t idx) {
    PCPO_ASSERT(idx < capacity, "schedule out of bound");
    if (!inside_queue[idx]) { inside_queue[idx] = 1; queue.push_back(idx); }
  }
  void unschedule(size_t idx) {
    PCPO_ASSERT(idx < capacity, "unschedule out of bound");
    if (inside_queue[idx]) {
      auto it = std::find(queue.begin(), queue.end(), idx);
      PCPO_ASSERT(it != queue.end(), "unschedule: not in queue");
      std::swap(*it, queue.front());  // VecDeque::swap_remove_front
      queue.pop_front();
      inside_queue[idx] = 0;
    }
  }
  bool pop(size_t* out) {
    if (queue.empty()) return false;
    *out = queue.front();
    queue.pop_front();
    inside_queue[*out] = 0;
    return true;
  }
  bool is_empty() const { return queue.empty(); }
};

}  // namespace pcpo
