// pcp_oracle_capi_body.hpp -- CPU ORACLE (test infrastructure): the engine behind the pcpo_* C
// entry points, compiled once per domain instantiation (pcpo::iv: Interval, pcpo::set:
// IntervalSet) by pcp_oracle_capi.cpp.  No include guard on purpose.

struct Engine : ::pcpo_engine {
  Space space;
  bool flat = false;
  std::vector<std::vector<pcp_operand>> sums;
  std::vector<Label> labels;

  Var make_view(pcp_operand op) const {
    if (op.var >= 0) {
      Var id = std::make_unique<Identity>(size_t(op.var));
      if (op.off == 0) return id;
      return std::make_unique<Addition>(std::move(id), op.off);
    }
    if (op.var == PCP_VAR_CONSTANT) return std::make_unique<Constant>(op.off);
    size_t sid = size_t(-2 - op.var);
    PCPO_ASSERT(sid < sums.size(), "unknown sum view");
    std::vector<Var> terms;
    for (auto& t : sums[sid]) terms.push_back(make_view(t));
    Var s = std::make_unique<Sum>(std::move(terms));
    if (op.off == 0) return s;
    return std::make_unique<Addition>(std::move(s), op.off);
  }
  void check_operand(pcp_operand op) const {
    if (op.var >= 0) PCPO_ASSERT(size_t(op.var) < space.vstore.size(), "operand variable out of range");
    else if (op.var <= -2) PCPO_ASSERT(size_t(-2 - op.var) < sums.size(), "unknown sum view");
  }

  // A formula tree in the prefix encoding of pcp_formula_alloc (include/pcp_b200.h): the
  // reference's boxed Formula<VStore> built node by node (logic/*.rs, propagators/cmp/*.rs).
  Formula build_formula(const int32_t*& p, const int32_t* end) const {
    PCPO_ASSERT(p < end, "truncated formula");
    const int32_t tag = *p++;
    auto operand = [&]() {
      PCPO_ASSERT(p + 2 <= end, "truncated formula");
      pcp_operand op{p[0], p[1]};
      p += 2;
      check_operand(op);
      return make_view(op);
    };
    switch (tag) {
      case PCP_F_CONJUNCTION:
      case PCP_F_DISJUNCTION: {
        PCPO_ASSERT(p < end && *p >= 1, "connective without children");
        const int32_t n = *p++;
        std::vector<Formula> fs;
        for (int32_t i = 0; i < n; ++i) fs.push_back(build_formula(p, end));
        if (tag == PCP_F_CONJUNCTION) return std::make_unique<Conjunction>(std::move(fs));
        return std::make_unique<Disjunction>(std::move(fs));
      }
      case PCP_F_BOOLEAN: { Var v = operand(); return std::make_unique<Boolean>(std::move(v)); }
      case PCP_F_BOOLEAN_NEG: { Var v = operand(); return std::make_unique<BooleanNeg>(Boolean(std::move(v))); }
      case PCP_F_NOT: { Formula f = build_formula(p, end); return f->not_(); }
      case PCP_F_LEAF + PCP_X_LESS_Y: { Var a = operand(), b = operand(); return std::make_unique<XLessY>(std::move(a), std::move(b)); }
      case PCP_F_LEAF + PCP_X_NEQ_Y: { Var a = operand(), b = operand(); return std::make_unique<XNeqY>(std::move(a), std::move(b)); }
      case PCP_F_LEAF + PCP_X_EQ_Y: { Var a = operand(), b = operand(); return std::make_unique<XEqY>(std::move(a), std::move(b)); }
      case PCP_F_LEAF + PCP_X_GREATER_Y_PLUS_Z: { Var a = operand(), b = operand(), c = operand(); return std::make_unique<XGreaterYPlusZ>(std::move(a), std::move(b), std::move(c)); }
      case PCP_F_LEAF + PCP_X_LESS_Y_PLUS_Z: { Var a = operand(), b = operand(), c = operand(); return std::make_unique<XLessYPlusZ>(std::move(a), std::move(b), std::move(c)); }
      case PCP_F_LEAF + PCP_X_EQ_Y_PLUS_Z: { Var a = operand(), b = operand(), c = operand(); return std::make_unique<XEqYPlusZ>(std::move(a), std::move(b), std::move(c)); }
      default: throw ContractViolation("unknown formula node");
    }
  }

  Formula make_prop(int kind, const pcp_operand* ops, int n) const {
    for (int i = 0; i < n; ++i) check_operand(ops[i]);
    if (flat) {  // same kind numbering as enum pcp_prop_kind
      std::vector<FOp> f;
      for (int i = 0; i < n; ++i) {
        pcp_operand op = ops[i];
        if (op.var <= -2) {  // single-term sums delegate (term/sum.rs:62-64); wider ones need the view tree
          const auto& terms = sums[size_t(-2 - op.var)];
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
    auto v = [&](int i) { return make_view(ops[i]); };
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

  void vars_alloc(const int32_t* lo, const int32_t* hi, int32_t n, int32_t* first_idx) override {
    for (int i = 0; i < n; ++i) PCPO_ASSERT(lo[i] <= hi[i], "alloc of an empty domain");
    if (first_idx) *first_idx = int32_t(space.vstore.size());
    for (int i = 0; i < n; ++i) space.vstore.alloc(Dom(lo[i], hi[i]));
  }
  void sum_alloc(const pcp_operand* terms, int32_t n, int32_t* sum_id) override {
    PCPO_ASSERT(n >= 1, "At least one variable in sum.");
    for (int i = 0; i < n; ++i) { PCPO_ASSERT(terms[i].var >= -1, "nested sums are not supported"); check_operand(terms[i]); }
    sums.emplace_back(terms, terms + n);
    if (sum_id) *sum_id = int32_t(sums.size() - 1);
  }
  void props_alloc(int32_t kind, const pcp_operand* ops, int32_t n_ops, int64_t n_props, int32_t* first_idx) override {
    if (first_idx) *first_idx = int32_t(space.cstore.size());
    for (int64_t p = 0; p < n_props; ++p) space.cstore.alloc(make_prop(kind, ops + p * n_ops, n_ops));
  }
  void formula_alloc(const int32_t* words, int32_t n_words, int32_t* idx) override {
    const int32_t* p = words;
    Formula f = build_formula(p, words + n_words);
    PCPO_ASSERT(p == words + n_words, "trailing words after the formula");
    size_t i = space.cstore.alloc(std::move(f));
    if (idx) *idx = int32_t(i);
  }
  void consistency(int32_t* status, pcp_stats* stats) override {
    uint64_t before = space.cstore.num_propagations;
    SKleene k = space.consistency();
    *status = int32_t(k);
    if (stats) {
      std::memset(stats, 0, sizeof(*stats));
      stats->propagations = space.cstore.num_propagations - before;
      uint32_t a = 0;
      for (uint8_t b : space.cstore.active) a += b;
      stats->active_props = a;
    }
  }
  void domains_read(int32_t first, int32_t n, int32_t* lo, int32_t* hi) override {
    PCPO_ASSERT(first >= 0 && n >= 0 && size_t(first) + size_t(n) <= space.vstore.size(), "Variable not registered in the store.");
    for (int i = 0; i < n; ++i) { lo[i] = space.vstore.memory[first + i].lower(); hi[i] = space.vstore.memory[first + i].upper(); }
  }
  void domains_size_read(int32_t first, int32_t n, uint32_t* size) override {
    PCPO_ASSERT(first >= 0 && n >= 0 && size_t(first) + size_t(n) <= space.vstore.size(), "Variable not registered in the store.");
    for (int i = 0; i < n; ++i) size[i] = space.vstore.memory[first + i].size();
  }
  void domains_read_bits(int32_t first, int32_t n, int32_t base, int32_t words, uint32_t* out) override {
    PCPO_ASSERT(first >= 0 && n >= 0 && words >= 0 && size_t(first) + size_t(n) <= space.vstore.size(), "Variable not registered in the store.");
    std::memset(out, 0, size_t(n) * size_t(words) * 4);
    for (int i = 0; i < n; ++i)
      space.vstore.memory[first + i].for_each_run([&](int32_t lb, int32_t ub) {
        for (int64_t v = lb; v <= ub; ++v) {
          int64_t b = v - base;
          PCPO_ASSERT(b >= 0 && b < int64_t(words) * 32, "value outside the requested bit window");
          out[size_t(i) * words + size_t(b >> 5)] |= 1u << (b & 31);
        }
      });
  }
  void var_update(int32_t idx, int32_t lo, int32_t hi, int32_t* ok) override {
    PCPO_ASSERT(idx >= 0 && size_t(idx) < space.vstore.size(), "Variable not registered in the store.");
    // MonotonicUpdate::update takes a domain; through the C ABI the new domain is `cur /\ [lo, hi]`
    // (on Interval domains that is [lo, hi] itself), with [lo, hi] inside the current bounds
    const Dom& cur = space.vstore.memory[size_t(idx)];
    PCPO_ASSERT(lo > hi || (lo >= cur.lower() && hi <= cur.upper()), "Domain update must be monotonic.");
    bool r = space.vstore.update(size_t(idx), lo > hi ? Dom::empty() : cur.shrink_left(lo).shrink_right(hi));
    space.vstore.drain_delta();  // the store starts every consistency() from all-scheduled
    if (ok) *ok = r ? 1 : 0;
  }
  void active_read(int32_t first, int32_t n, uint8_t* out) override {
    PCPO_ASSERT(first >= 0 && n >= 0 && size_t(first) + size_t(n) <= space.cstore.size(), "propagator out of range");
    for (int i = 0; i < n; ++i) out[i] = space.cstore.active[first + i];
  }
  void label(uint64_t* l) override { labels.push_back(space.label()); *l = labels.size() - 1; }
  void restore(uint64_t l) override {
    PCPO_ASSERT(l < labels.size(), "unknown label");
    space.restore(labels[l]);
    labels.resize(l + 1);
  }
  int32_t num_vars() const override { return int32_t(space.vstore.size()); }
  int32_t num_props() const override { return int32_t(space.cstore.size()); }

  // The reference's propagator test fixture (propagators/mod.rs:110-129):
  // is_subsumed before; propagate; ordered delta; is_subsumed after.
  void test_propagation(int32_t prop, int32_t* before, int32_t* propagate_ok, int32_t* after, int32_t* delta,
                        int32_t* n_delta) override {
    PCPO_ASSERT(prop >= 0 && size_t(prop) < space.cstore.size(), "propagator out of range");
    Propagator& p = *space.cstore.propagators[prop];
    VStore& vs = space.vstore;
    *before = int32_t(p.is_subsumed(vs));
    bool ok = p.propagate(vs);
    *propagate_ok = ok ? 1 : 0;
    auto d = vs.drain_delta();
    int cap = *n_delta;
    *n_delta = int32_t(d.size());
    for (int i = 0; i < int(d.size()) && i < cap; ++i) { delta[2 * i] = int32_t(d[i].first); delta[2 * i + 1] = int32_t(d[i].second); }
    *after = int32_t(p.is_subsumed(vs));
  }
  // PropagatorDependencies::dependencies (propagation/ops.rs:27-29).
  void prop_dependencies(int32_t prop, int32_t* deps, int32_t* n_deps) override {
    PCPO_ASSERT(prop >= 0 && size_t(prop) < space.cstore.size(), "propagator out of range");
    auto d = space.cstore.propagators[prop]->dependencies();
    int cap = *n_deps;
    *n_deps = int32_t(d.size());
    for (int i = 0; i < int(d.size()) && i < cap; ++i) { deps[2 * i] = int32_t(d[i].first); deps[2 * i + 1] = int32_t(d[i].second); }
  }
  // Store::is_subsumed of the whole constraint store (propagation/store.rs: Kleene-and of the
  // active propagators), as cumulative.rs:236-252 asserts before and after consistency().
  int32_t store_is_subsumed() const override { return int32_t(space.cstore.is_subsumed(space.vstore)); }

  void search_run(const pcp_search_config* cfg, pcp_search_result* res, int32_t* trace_status, uint64_t* trace_hash,
                  int32_t* trace_lo, int32_t* trace_hi, uint64_t trace_capacity) override {
    Search s;
    s.flat = flat;
    s.cfg.node_limit = cfg->node_limit;
    s.cfg.all_solutions = cfg->all_solutions != 0;
    s.cfg.var_sel = cfg->var_sel;
    s.cfg.val_sel = cfg->val_sel;
    s.cfg.distributor = cfg->distributor;
    s.cfg.bb_mode = BBMode(cfg->bb_mode);
    s.cfg.bb_var = size_t(cfg->bb_var);
    uint64_t n = 0;
    size_t V = space.vstore.size();
    uint64_t p_warm = space.cstore.num_propagations;
    auto t0 = std::chrono::steady_clock::now();
    s.on_node = [&](const Space& sp, int status) {
      if (n + 1 == uint64_t(cfg->warmup_nodes)) { t0 = std::chrono::steady_clock::now(); p_warm = space.cstore.num_propagations; }
      if (n < trace_capacity) {
        if (trace_status) trace_status[n] = status;
        if (trace_hash) trace_hash[n] = status == int(False) ? 0 : hash_domains(sp.vstore.memory);
        if (cfg->trace_domains && trace_lo && trace_hi && status != int(False))
          for (size_t i = 0; i < V; ++i) { trace_lo[n * V + i] = sp.vstore.memory[i].lower(); trace_hi[n * V + i] = sp.vstore.memory[i].upper(); }
      }
      ++n;
    };
    NodeStatus st = s.run(space);
    auto t1 = std::chrono::steady_clock::now();
    std::memset(res, 0, sizeof(*res));
    res->status = int32_t(st);
    res->has_bb_value = s.has_bb_value;
    res->bb_value = s.bb_value;
    res->num_nodes = s.stats.num_nodes;
    res->num_solution = s.stats.num_solution;
    res->num_failed_node = s.stats.num_failed_node;
    res->num_prune = s.stats.num_prune;
    res->propagations = space.cstore.num_propagations - p_warm;
    res->seconds = std::chrono::duration<double>(t1 - t0).count();
  }
};
