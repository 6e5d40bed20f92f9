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
#pragma once
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
#include "pcp_oracle_common.hpp"

namespace pcpo {
namespace iv {   // libpcp instantiated at VStoreFD = VStoreTrail<Interval<i32>> (variable/mod.rs:37)
using Dom = Interval;
#include "pcp_oracle_body.hpp"
}  // namespace iv
namespace set {  // libpcp instantiated at VStoreSet = VStoreTrail<IntervalSet<i32>> (variable/mod.rs:38): FDSpace
using Dom = IntervalSet;
#include "pcp_oracle_body.hpp"
}  // namespace set
}  // namespace pcpo
