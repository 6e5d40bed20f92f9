// pcp_oracle_body.hpp -- CPU ORACLE (test infrastructure): the part of the restatement that is
// generic over the domain type, included once per instantiation by pcp_oracle.hpp with `Dom`
// defined: pcpo::iv (Dom = Interval, libpcp's VStoreFD) and pcpo::set (Dom = IntervalSet,
// libpcp's VStoreSet / FDSpace).  No include guard on purpose.

// events/mod.rs:46-70; returns -1 for None.
inline int event_new(const Dom& little, const Dom& big) {
  PCPO_ASSERT(little.is_subset(big), "Events are computed on the difference between `little` and `big`.");
  if (little.size() != big.size()) {
    if (little.is_singleton()) return Assignment;
    if (little.lower() != big.lower() || little.upper() != big.upper()) return Bound;
    return Inner;
  }
  return -1;
}

// ---------------------------------------------------------------------------
// variable/store.rs -- Store<Memory,Event> with delta + has_changed.
// Memory here is CopyMemory (variable/memory/copy_memory.rs:115-152): labels
// are shared snapshots of the whole domain vector.
// ---------------------------------------------------------------------------
struct VStore {
  std::vector<Dom> memory;
  std::vector<int8_t> delta_ev;      // VecMap<Event>: -1 = absent
  std::vector<uint32_t> delta_keys;  // inserted keys (drained in ascending order)
  bool changed = false;

  size_t size() const { return memory.size(); }
  const Dom& operator[](size_t i) const {  // store.rs:175-181
    PCPO_ASSERT(i < memory.size(), "Variable not registered in the store.");
    return memory[i];
  }
  size_t alloc(const Dom& d) {  // store.rs:135-140
    PCPO_ASSERT(!d.is_empty(), "alloc of an empty domain");
    memory.push_back(d);
    delta_ev.push_back(-1);
    return memory.size() - 1;
  }
  void update_delta(size_t key, const Dom& old_dom) {  // store.rs:94-106
    int ev = event_new(memory[key], old_dom);
    if (ev >= 0) {
      changed = true;
      if (delta_ev[key] >= 0) delta_ev[key] = int8_t(merge(FDEvent(delta_ev[key]), FDEvent(ev)));
      else { delta_ev[key] = int8_t(ev); delta_keys.push_back(uint32_t(key)); }
    }
  }
  bool update(size_t idx, const Dom& dom) {  // store.rs:151-166
    PCPO_ASSERT(idx < memory.size(), "Variable not registered in the store.");
    PCPO_ASSERT(dom.is_subset(memory[idx]), "Domain update must be monotonic.");
    if (dom.is_empty()) return false;
    if (dom.size() < memory[idx].size()) {
      Dom old = memory[idx];
      memory[idx] = dom;
      update_delta(idx, old);
    }
    return true;
  }
  // store.rs:225-237 (vec_map drain = ascending key order).
  std::vector<Dep> drain_delta() {
    std::sort(delta_keys.begin(), delta_keys.end());
    std::vector<Dep> out;
    out.reserve(delta_keys.size());
    for (uint32_t k : delta_keys) { out.emplace_back(k, FDEvent(delta_ev[k])); delta_ev[k] = -1; }
    delta_keys.clear();
    return out;
  }
  bool has_changed() const { return changed; }
  void reset_changed() { changed = false; }
};

// ---------------------------------------------------------------------------
// term/ -- views.  Var<VStore> = Box<dyn IntVariable> (concept.rs:79-118).
// ---------------------------------------------------------------------------
struct View {
  virtual ~View() = default;
  virtual Dom read(const VStore&) const = 0;
  virtual bool update(VStore&, const Dom&) = 0;
  virtual std::vector<Dep> dependencies(FDEvent) const = 0;
  virtual std::unique_ptr<View> bclone() const = 0;
};
using Var = std::unique_ptr<View>;

struct Identity : View {  // term/identity.rs:47-70
  size_t idx;
  explicit Identity(size_t i) : idx(i) {}
  Dom read(const VStore& s) const override { return s[idx]; }
  bool update(VStore& s, const Dom& v) override { return s.update(idx, v); }
  std::vector<Dep> dependencies(FDEvent ev) const override { return {{idx, ev}}; }
  Var bclone() const override { return std::make_unique<Identity>(idx); }
};
struct Addition : View {  // term/addition.rs:80-110
  Var x; int32_t v;
  Addition(Var x_, int32_t v_) : x(std::move(x_)), v(v_) {}
  Dom read(const VStore& s) const override { return x->read(s).plus(v); }
  bool update(VStore& s, const Dom& value) override { return x->update(s, value.minus(v)); }
  std::vector<Dep> dependencies(FDEvent ev) const override { return x->dependencies(ev); }
  Var bclone() const override { return std::make_unique<Addition>(x->bclone(), v); }
};
struct Constant : View {  // term/constant.rs:43-68
  int32_t value;
  explicit Constant(int32_t v) : value(v) {}
  Dom read(const VStore&) const override { return Dom::singleton(value); }
  bool update(VStore&, const Dom& v) override { return !v.is_empty() && v.contains(value); }
  std::vector<Dep> dependencies(FDEvent) const override { return {}; }
  Var bclone() const override { return std::make_unique<Constant>(value); }
};
struct Sum : View {  // term/sum.rs:56-92
  std::vector<Var> vars;
  explicit Sum(std::vector<Var> v) : vars(std::move(v)) {}
  Dom read(const VStore& s) const override {
    PCPO_ASSERT(!vars.empty(), "At least one variable in sum.");
    Dom a = vars[0]->read(s);
    for (size_t i = 1; i < vars.size(); ++i) a = a.add(vars[i]->read(s));
    return a;
  }
  bool update(VStore& s, const Dom& value) override {
    if (vars.size() == 1) return vars[0]->update(s, value);
    return read(s).overlap(value);
  }
  std::vector<Dep> dependencies(FDEvent ev) const override {
    std::vector<Dep> d;
    for (auto& v : vars) { auto dv = v->dependencies(ev); d.insert(d.end(), dv.begin(), dv.end()); }
    return d;
  }
  Var bclone() const override {
    std::vector<Var> c; for (auto& v : vars) c.push_back(v->bclone());
    return std::make_unique<Sum>(std::move(c));
  }
};

// ---------------------------------------------------------------------------
// propagation/ops.rs:17-29, propagation/concept.rs:21-53 -- propagator surface.
// ---------------------------------------------------------------------------
struct Propagator {
  virtual ~Propagator() = default;
  virtual bool propagate(VStore&) = 0;
  virtual SKleene is_subsumed(const VStore&) const = 0;
  virtual std::vector<Dep> dependencies() const = 0;
  virtual std::unique_ptr<Propagator> bclone() const = 0;
  virtual std::unique_ptr<Propagator> not_() const { throw ContractViolation("not(): unimplemented"); }
};
using Formula = std::unique_ptr<Propagator>;

inline std::vector<Dep> cat(std::vector<Dep> a, const std::vector<Dep>& b) {
  a.insert(a.end(), b.begin(), b.end());
  return a;
}

struct XLessY;
Formula make_x_geq_y(Var x, Var y);
Formula make_x_leq_y_plus_z(Var x, Var y, Var z);
Formula make_x_geq_y_plus_z(Var x, Var y, Var z);

// propagators/cmp/x_less_y.rs:67-117
struct XLessY : Propagator {
  Var x, y;
  XLessY(Var x_, Var y_) : x(std::move(x_)), y(std::move(y_)) {}
  SKleene is_subsumed(const VStore& s) const override {
    Dom a = x->read(s), b = y->read(s);
    if (a.lower() >= b.upper()) return False;
    if (a.upper() < b.lower()) return True;
    return Unknown;
  }
  bool propagate(VStore& s) override {
    Dom a = x->read(s), b = y->read(s);
    return x->update(s, a.strict_shrink_right(b.upper())) &&
           y->update(s, b.strict_shrink_left(a.lower()));
  }
  std::vector<Dep> dependencies() const override {
    return cat(x->dependencies(Bound), y->dependencies(Bound));
  }
  Formula bclone() const override { return std::make_unique<XLessY>(x->bclone(), y->bclone()); }
  Formula not_() const override { return make_x_geq_y(x->bclone(), y->bclone()); }
};

struct XNeqY;
// propagators/cmp/x_eq_y.rs:67-116
struct XEqY : Propagator {
  Var x, y;
  XEqY(Var x_, Var y_) : x(std::move(x_)), y(std::move(y_)) {}
  SKleene is_subsumed(const VStore& s) const override {
    Dom a = x->read(s), b = y->read(s);
    if (a.lower() == b.upper() && a.upper() == b.lower()) return True;
    if (a.is_disjoint(b)) return False;
    return Unknown;
  }
  bool propagate(VStore& s) override {
    Dom a = x->read(s), b = y->read(s);
    Dom n = a.intersection(b);
    return x->update(s, n) && y->update(s, n);
  }
  std::vector<Dep> dependencies() const override {
    return cat(x->dependencies(Inner), y->dependencies(Inner));
  }
  Formula bclone() const override { return std::make_unique<XEqY>(x->bclone(), y->bclone()); }
  Formula not_() const override;
};

// propagators/cmp/x_neq_y.rs:66-104
struct XNeqY : Propagator {
  Var x, y;
  XNeqY(Var x_, Var y_) : x(std::move(x_)), y(std::move(y_)) {}
  SKleene is_subsumed(const VStore& s) const override {
    return k_not(XEqY(x->bclone(), y->bclone()).is_subsumed(s));
  }
  bool propagate(VStore& s) override {
    Dom a = x->read(s), b = y->read(s);
    if (a.is_singleton()) return y->update(s, b.difference(a.lower()));
    if (b.is_singleton()) return x->update(s, a.difference(b.lower()));
    return true;
  }
  std::vector<Dep> dependencies() const override {
    return XEqY(x->bclone(), y->bclone()).dependencies();
  }
  Formula bclone() const override { return std::make_unique<XNeqY>(x->bclone(), y->bclone()); }
  Formula not_() const override { return std::make_unique<XEqY>(x->bclone(), y->bclone()); }
};
inline Formula XEqY::not_() const { return std::make_unique<XNeqY>(x->bclone(), y->bclone()); }

// propagators/cmp/x_greater_y_plus_z.rs:75-128
struct XGreaterYPlusZ : Propagator {
  Var x, y, z;
  XGreaterYPlusZ(Var x_, Var y_, Var z_) : x(std::move(x_)), y(std::move(y_)), z(std::move(z_)) {}
  SKleene is_subsumed(const VStore& s) const override {
    Dom a = x->read(s), b = y->read(s), c = z->read(s);
    if (a.upper() <= b.lower() + c.lower()) return False;
    if (a.lower() > b.upper() + c.upper()) return True;
    return Unknown;
  }
  bool propagate(VStore& s) override {
    Dom a = x->read(s), b = y->read(s), c = z->read(s);
    return x->update(s, a.strict_shrink_left(b.lower() + c.lower())) &&
           y->update(s, b.strict_shrink_right(a.upper() - c.lower())) &&
           z->update(s, c.strict_shrink_right(a.upper() - b.lower()));
  }
  std::vector<Dep> dependencies() const override {
    return cat(cat(x->dependencies(Bound), y->dependencies(Bound)), z->dependencies(Bound));
  }
  Formula bclone() const override {
    return std::make_unique<XGreaterYPlusZ>(x->bclone(), y->bclone(), z->bclone());
  }
  Formula not_() const override { return make_x_leq_y_plus_z(x->bclone(), y->bclone(), z->bclone()); }
};

// propagators/cmp/x_less_y_plus_z.rs:75-128
struct XLessYPlusZ : Propagator {
  Var x, y, z;
  XLessYPlusZ(Var x_, Var y_, Var z_) : x(std::move(x_)), y(std::move(y_)), z(std::move(z_)) {}
  SKleene is_subsumed(const VStore& s) const override {
    Dom a = x->read(s), b = y->read(s), c = z->read(s);
    if (a.lower() >= b.upper() + c.upper()) return False;
    if (a.upper() < b.lower() + c.lower()) return True;
    return Unknown;
  }
  bool propagate(VStore& s) override {
    Dom a = x->read(s), b = y->read(s), c = z->read(s);
    return x->update(s, a.strict_shrink_right(b.upper() + c.upper())) &&
           y->update(s, b.strict_shrink_left(a.lower() - c.upper())) &&
           z->update(s, c.strict_shrink_left(a.lower() - b.upper()));
  }
  std::vector<Dep> dependencies() const override {
    return cat(cat(x->dependencies(Bound), y->dependencies(Bound)), z->dependencies(Bound));
  }
  Formula bclone() const override {
    return std::make_unique<XLessYPlusZ>(x->bclone(), y->bclone(), z->bclone());
  }
  Formula not_() const override { return make_x_geq_y_plus_z(x->bclone(), y->bclone(), z->bclone()); }
};

// propagators/cmp/mod.rs:40-86 -- derived constructors.
inline Formula make_x_greater_y(Var x, Var y) { return std::make_unique<XLessY>(std::move(y), std::move(x)); }
inline Formula make_x_geq_y(Var x, Var y) {
  return make_x_greater_y(std::make_unique<Addition>(std::move(x), 1), std::move(y));
}
inline Formula make_x_leq_y(Var x, Var y) {
  return std::make_unique<XLessY>(std::move(x), std::make_unique<Addition>(std::move(y), 1));
}
inline Formula make_x_geq_y_plus_z(Var x, Var y, Var z) {
  return std::make_unique<XGreaterYPlusZ>(std::make_unique<Addition>(std::move(x), 1), std::move(y), std::move(z));
}
inline Formula make_x_leq_y_plus_z(Var x, Var y, Var z) {
  return std::make_unique<XLessYPlusZ>(std::make_unique<Addition>(std::move(x), -1), std::move(y), std::move(z));
}

// propagators/cmp/x_eq_y_plus_z.rs:26-105
struct XEqYPlusZ : Propagator {
  Formula geq, leq;
  XEqYPlusZ(Var x, Var y, Var z) {
    geq = make_x_geq_y_plus_z(x->bclone(), y->bclone(), z->bclone());
    leq = make_x_leq_y_plus_z(std::move(x), std::move(y), std::move(z));
  }
  XEqYPlusZ(Formula g, Formula l) : geq(std::move(g)), leq(std::move(l)) {}
  SKleene is_subsumed(const VStore& s) const override {
    return k_and(geq->is_subsumed(s), leq->is_subsumed(s));
  }
  bool propagate(VStore& s) override { return geq->propagate(s) && leq->propagate(s); }
  std::vector<Dep> dependencies() const override {
    auto g = geq->dependencies(), l = leq->dependencies();
    PCPO_ASSERT(g == l, "dependencies of X >= Y + Z and X <= Y + Z differ");
    return g;
  }
  Formula bclone() const override { return std::make_unique<XEqYPlusZ>(geq->bclone(), leq->bclone()); }
};

// propagators/cmp/x_eq_y_mul_z.rs:68-116.
struct XEqYMulZ : Propagator {
  Var x, y, z;
  XEqYMulZ(Var x_, Var y_, Var z_) : x(std::move(x_)), y(std::move(y_)), z(std::move(z_)) {}
  SKleene is_subsumed(const VStore& s) const override {
    Dom a = x->read(s), b = y->read(s), c = z->read(s);
    // x_eq_y_mul_z.rs:73-90: overlap -> True iff y*z and x are both singletons, else Unknown;
    // no overlap -> False (the product of two intervals can be a singleton without y and z
    // being singletons: [0,0] * [1,5])
    Dom yz = b.mul(c);
    if (a.is_disjoint(yz)) return False;
    return (yz.is_singleton() && a.is_singleton()) ? True : Unknown;
  }
  bool propagate(VStore& s) override {
    Dom a = x->read(s), b = y->read(s), c = z->read(s);
    return x->update(s, a.intersection(b.mul(c)));
  }
  std::vector<Dep> dependencies() const override {
    return cat(cat(x->dependencies(Bound), y->dependencies(Bound)), z->dependencies(Bound));
  }
  Formula bclone() const override { return std::make_unique<XEqYMulZ>(x->bclone(), y->bclone(), z->bclone()); }
};

inline std::vector<Dep> sorted_dedup(std::vector<Dep> d) {
  std::sort(d.begin(), d.end());
  d.erase(std::unique(d.begin(), d.end()), d.end());
  return d;
}

struct Disjunction;
// logic/conjunction.rs:77-118
struct Conjunction : Propagator {
  std::vector<Formula> fs;
  explicit Conjunction(std::vector<Formula> f) : fs(std::move(f)) {}
  SKleene is_subsumed(const VStore& s) const override {
    bool all_entailed = true;
    for (auto& f : fs) {
      SKleene k = f->is_subsumed(s);
      if (k == False) return False;
      if (k == Unknown) all_entailed = false;
    }
    return all_entailed ? True : Unknown;
  }
  bool propagate(VStore& s) override {
    for (auto& f : fs) if (!f->propagate(s)) return false;
    return true;
  }
  std::vector<Dep> dependencies() const override {
    std::vector<Dep> d;
    for (auto& f : fs) d = cat(std::move(d), f->dependencies());
    return sorted_dedup(std::move(d));
  }
  Formula bclone() const override {
    std::vector<Formula> c; for (auto& f : fs) c.push_back(f->bclone());
    return std::make_unique<Conjunction>(std::move(c));
  }
  Formula not_() const override;
};

// logic/disjunction.rs:77-129
struct Disjunction : Propagator {
  std::vector<Formula> fs;
  explicit Disjunction(std::vector<Formula> f) : fs(std::move(f)) {}
  SKleene is_subsumed(const VStore& s) const override {
    bool all_disentailed = true;
    for (auto& f : fs) {
      SKleene k = f->is_subsumed(s);
      if (k == True) return True;
      if (k == Unknown) all_disentailed = false;
    }
    return all_disentailed ? False : Unknown;
  }
  bool propagate(VStore& s) override {
    size_t num_disentailed = 0, unknown_formula = 0;
    for (size_t i = 0; i < fs.size(); ++i) {
      SKleene k = fs[i]->is_subsumed(s);
      if (k == True) return true;
      if (k == False) ++num_disentailed; else unknown_formula = i;
    }
    if (num_disentailed == fs.size() - 1) return fs[unknown_formula]->propagate(s);
    if (num_disentailed == fs.size()) return false;
    return true;
  }
  std::vector<Dep> dependencies() const override {
    std::vector<Dep> d;
    for (auto& f : fs) d = cat(std::move(d), f->dependencies());
    return sorted_dedup(std::move(d));
  }
  Formula bclone() const override {
    std::vector<Formula> c; for (auto& f : fs) c.push_back(f->bclone());
    return std::make_unique<Disjunction>(std::move(c));
  }
  Formula not_() const override {
    std::vector<Formula> c; for (auto& f : fs) c.push_back(f->not_());
    return std::make_unique<Conjunction>(std::move(c));
  }
};
inline Formula Conjunction::not_() const {
  std::vector<Formula> c; for (auto& f : fs) c.push_back(f->not_());
  return std::make_unique<Disjunction>(std::move(c));
}

// logic/boolean.rs:29-141 -- a 0/1 variable read as a formula.  Boolean::new allocates the
// variable ([0, 1], boolean.rs:35-40); here the caller passes the view.
struct BooleanNeg;
struct Boolean : Propagator {
  Var var;
  explicit Boolean(Var v) : var(std::move(v)) {}
  SKleene is_subsumed(const VStore& s) const override {  // boolean.rs:108-122
    Dom x = var->read(s);
    if (x.is_singleton()) return x.lower() == 1 ? True : False;
    return Unknown;
  }
  bool propagate(VStore& s) override { return var->update(s, Dom::singleton(1)); }  // boolean.rs:126-135
  std::vector<Dep> dependencies() const override { return var->dependencies(Bound); }  // boolean.rs:137-141
  Formula bclone() const override { return std::make_unique<Boolean>(var->bclone()); }
  Formula not_() const override;  // boolean.rs:70-79
};
// logic/boolean_neg.rs:30-98
struct BooleanNeg : Propagator {
  Boolean b;
  explicit BooleanNeg(Boolean b_) : b(std::move(b_)) {}
  SKleene is_subsumed(const VStore& s) const override { return k_not(b.is_subsumed(s)); }
  bool propagate(VStore& s) override { return b.var->update(s, Dom::singleton(0)); }
  std::vector<Dep> dependencies() const override { return b.dependencies(); }
  Formula bclone() const override { return std::make_unique<BooleanNeg>(Boolean(b.var->bclone())); }
  Formula not_() const override { return std::make_unique<Boolean>(b.var->bclone()); }
};
inline Formula Boolean::not_() const { return std::make_unique<BooleanNeg>(Boolean(var->bclone())); }

// logic/mod.rs:30-45 (note the operand order of `implication`: Disjunction[f, not g])
inline Formula make_implication(Formula f, Formula g) {
  std::vector<Formula> fs;
  Formula ng = g->not_();
  fs.push_back(std::move(f));
  fs.push_back(std::move(ng));
  return std::make_unique<Disjunction>(std::move(fs));
}
inline Formula make_equivalence(Formula f, Formula g) {
  std::vector<Formula> fs;
  fs.push_back(make_implication(f->bclone(), g->bclone()));
  fs.push_back(make_implication(std::move(g), std::move(f)));
  return std::make_unique<Conjunction>(std::move(fs));
}

// propagators/distinct.rs:69-126
struct Distinct : Propagator {
  Conjunction conj;
  std::vector<Var> vars;
  static std::vector<Formula> pairs(const std::vector<Var>& v) {
    PCPO_ASSERT(!v.empty(), "Variable array in `Distinct` must be non-empty.");
    std::vector<Formula> props;
    for (size_t i = 0; i + 1 < v.size(); ++i)
      for (size_t j = i + 1; j < v.size(); ++j)
        props.push_back(std::make_unique<XNeqY>(v[i]->bclone(), v[j]->bclone()));
    return props;
  }
  explicit Distinct(std::vector<Var> v) : conj(pairs(v)), vars(std::move(v)) {}
  SKleene is_subsumed(const VStore& s) const override { return conj.is_subsumed(s); }
  bool propagate(VStore& s) override { return conj.propagate(s); }
  std::vector<Dep> dependencies() const override {
    std::vector<Dep> d;
    for (auto& v : vars) d = cat(std::move(d), v->dependencies(Inner));
    return d;
  }
  Formula bclone() const override {
    std::vector<Var> c; for (auto& v : vars) c.push_back(v->bclone());
    return std::make_unique<Distinct>(std::move(c));
  }
  Formula not_() const override { return conj.not_(); }
};

// propagators/all_equal.rs:47-103: chain of XEqY(vars[i], vars[i+1]); dependencies are the
// variables themselves at Inner, each once (all_equal.rs:96-103).
struct AllEqual : Propagator {
  Conjunction conj;
  std::vector<Var> vars;
  static std::vector<Formula> chain(const std::vector<Var>& v) {
    PCPO_ASSERT(!v.empty(), "Variable array in `AllEqual` must be non-empty.");
    std::vector<Formula> props;
    for (size_t i = 0; i + 1 < v.size(); ++i)
      props.push_back(std::make_unique<XEqY>(v[i]->bclone(), v[i + 1]->bclone()));
    return props;
  }
  explicit AllEqual(std::vector<Var> v) : conj(chain(v)), vars(std::move(v)) {}
  SKleene is_subsumed(const VStore& s) const override { return conj.is_subsumed(s); }
  bool propagate(VStore& s) override { return conj.propagate(s); }
  std::vector<Dep> dependencies() const override {
    std::vector<Dep> d;
    for (auto& v : vars) d = cat(std::move(d), v->dependencies(Inner));
    return d;
  }
  Formula bclone() const override {
    std::vector<Var> c; for (auto& v : vars) c.push_back(v->bclone());
    return std::make_unique<AllEqual>(std::move(c));
  }
};

// ---------------------------------------------------------------------------
// FlatProp -- the same propagators without the Box<dyn ..> trees: operands are
// (var, off) pairs held inline, no heap allocation or view dispatch per call.
// Semantically identical to the classes above (update order, short-circuit,
// events); it exists so that the timed CPU baseline is a *tuned* implementation
// rather than one that pays the reference's allocation pattern
// (x_neq_y.rs:72,102 bclone()s two views per call).  tests/ check that it
// produces the same traces as the faithful restatement.
// ---------------------------------------------------------------------------
struct FOp { int32_t var, off; };  // var >= 0: Identity+off; var == -1: Constant(off)

struct FlatProp : Propagator {
  enum Kind : int { LessY = 0, NeqY = 1, EqY = 2, GreaterYPlusZ = 3, LessYPlusZ = 4, EqYPlusZ = 5,
                    DistinctN = 6, Disj2EqYPlusZ = 7, EqYMulZ = 8, AllEqualN = 9 };
  int kind;
  FOp o[6];
  std::vector<FOp> nary;

  static Dom rd(const VStore& s, FOp a) {
    return a.var >= 0 ? s.memory[size_t(a.var)].plus(a.off) : Dom::singleton(a.off);
  }
  static bool up(VStore& s, FOp a, const Dom& v) {
    if (a.var >= 0) return s.update(size_t(a.var), v.minus(a.off));
    return !v.is_empty() && v.contains(a.off);
  }
  static void dep(std::vector<Dep>& d, FOp a, FDEvent ev) { if (a.var >= 0) d.emplace_back(size_t(a.var), ev); }

  static SKleene sub_less(const VStore& s, FOp x, FOp y) {
    Dom a = rd(s, x), b = rd(s, y);
    if (a.lower() >= b.upper()) return False;
    if (a.upper() < b.lower()) return True;
    return Unknown;
  }
  static SKleene sub_eq(const VStore& s, FOp x, FOp y) {
    Dom a = rd(s, x), b = rd(s, y);
    if (a.lower() == b.upper() && a.upper() == b.lower()) return True;
    if (a.is_disjoint(b)) return False;
    return Unknown;
  }
  static bool prop_neq(VStore& s, FOp x, FOp y) {
    Dom a = rd(s, x), b = rd(s, y);
    if (a.is_singleton()) return up(s, y, b.difference(a.lower()));
    if (b.is_singleton()) return up(s, x, a.difference(b.lower()));
    return true;
  }
  // x (+xs) > y + z  /  x (+xs) < y + z with the Addition(x, +-1) of cmp/mod.rs:62-86 folded in xs
  static SKleene sub_greater(const VStore& s, FOp x, FOp y, FOp z, int xs) {
    Dom a = rd(s, x).plus(xs), b = rd(s, y), c = rd(s, z);
    if (a.upper() <= b.lower() + c.lower()) return False;
    if (a.lower() > b.upper() + c.upper()) return True;
    return Unknown;
  }
  static SKleene sub_lessyz(const VStore& s, FOp x, FOp y, FOp z, int xs) {
    Dom a = rd(s, x).plus(xs), b = rd(s, y), c = rd(s, z);
    if (a.lower() >= b.upper() + c.upper()) return False;
    if (a.upper() < b.lower() + c.lower()) return True;
    return Unknown;
  }
  static bool prop_greater(VStore& s, FOp x, FOp y, FOp z, int xs) {
    FOp xx{x.var, x.off + xs};
    Dom a = rd(s, xx), b = rd(s, y), c = rd(s, z);
    return up(s, xx, a.strict_shrink_left(b.lower() + c.lower())) && up(s, y, b.strict_shrink_right(a.upper() - c.lower())) &&
           up(s, z, c.strict_shrink_right(a.upper() - b.lower()));
  }
  static bool prop_lessyz(VStore& s, FOp x, FOp y, FOp z, int xs) {
    FOp xx{x.var, x.off + xs};
    Dom a = rd(s, xx), b = rd(s, y), c = rd(s, z);
    return up(s, xx, a.strict_shrink_right(b.upper() + c.upper())) && up(s, y, b.strict_shrink_left(a.lower() - c.upper())) &&
           up(s, z, c.strict_shrink_left(a.lower() - b.upper()));
  }
  static SKleene sub_eqyz(const VStore& s, const FOp* t) {
    return k_and(sub_greater(s, t[0], t[1], t[2], 1), sub_lessyz(s, t[0], t[1], t[2], -1));
  }
  static bool prop_eqyz(VStore& s, const FOp* t) {
    return prop_greater(s, t[0], t[1], t[2], 1) && prop_lessyz(s, t[0], t[1], t[2], -1);
  }

  SKleene is_subsumed(const VStore& s) const override {
    switch (kind) {
      case LessY: return sub_less(s, o[0], o[1]);
      case NeqY: return k_not(sub_eq(s, o[0], o[1]));
      case EqY: return sub_eq(s, o[0], o[1]);
      case GreaterYPlusZ: return sub_greater(s, o[0], o[1], o[2], 0);
      case LessYPlusZ: return sub_lessyz(s, o[0], o[1], o[2], 0);
      case EqYPlusZ: return sub_eqyz(s, o);
      case EqYMulZ: {  // x_eq_y_mul_z.rs:73-90
        Dom a = rd(s, o[0]), yz = rd(s, o[1]).mul(rd(s, o[2]));
        if (a.is_disjoint(yz)) return False;
        return (yz.is_singleton() && a.is_singleton()) ? True : Unknown;
      }
      case AllEqualN: {  // Conjunction::is_subsumed over the chain (all_equal.rs:53-57, conjunction.rs:77-94)
        bool all = true;
        for (size_t i = 0; i + 1 < nary.size(); ++i) {
          SKleene k = sub_eq(s, nary[i], nary[i + 1]);
          if (k == False) return False;
          if (k == Unknown) all = false;
        }
        return all ? True : Unknown;
      }
      case DistinctN: {  // Conjunction::is_subsumed over the pairs (conjunction.rs:77-94)
        bool all = true;
        for (size_t i = 0; i + 1 < nary.size(); ++i)
          for (size_t j = i + 1; j < nary.size(); ++j) {
            SKleene k = k_not(sub_eq(s, nary[i], nary[j]));
            if (k == False) return False;
            if (k == Unknown) all = false;
          }
        return all ? True : Unknown;
      }
      default: {  // Disjunction::is_subsumed (disjunction.rs:77-94)
        SKleene a = sub_eqyz(s, o), b = sub_eqyz(s, o + 3);
        if (a == True || b == True) return True;
        return (a == False && b == False) ? False : Unknown;
      }
    }
  }
  bool propagate(VStore& s) override {
    switch (kind) {
      case LessY: {
        Dom a = rd(s, o[0]), b = rd(s, o[1]);
        return up(s, o[0], a.strict_shrink_right(b.upper())) && up(s, o[1], b.strict_shrink_left(a.lower()));
      }
      case NeqY: return prop_neq(s, o[0], o[1]);
      case EqY: {
        Dom a = rd(s, o[0]), b = rd(s, o[1]);
        Dom n = a.intersection(b);
        return up(s, o[0], n) && up(s, o[1], n);
      }
      case GreaterYPlusZ: return prop_greater(s, o[0], o[1], o[2], 0);
      case LessYPlusZ: return prop_lessyz(s, o[0], o[1], o[2], 0);
      case EqYPlusZ: return prop_eqyz(s, o);
      case EqYMulZ: {  // x_eq_y_mul_z.rs:99-105
        Dom a = rd(s, o[0]);
        return up(s, o[0], a.intersection(rd(s, o[1]).mul(rd(s, o[2]))));
      }
      case AllEqualN:  // Conjunction::propagate over the chain of XEqY (x_eq_y.rs:102-107)
        for (size_t i = 0; i + 1 < nary.size(); ++i) {
          Dom a = rd(s, nary[i]), b = rd(s, nary[i + 1]);
          Dom n = a.intersection(b);
          if (!(up(s, nary[i], n) && up(s, nary[i + 1], n))) return false;
        }
        return true;
      case DistinctN:  // Conjunction::propagate over the pairs (conjunction.rs:96-105)
        for (size_t i = 0; i + 1 < nary.size(); ++i)
          for (size_t j = i + 1; j < nary.size(); ++j)
            if (!prop_neq(s, nary[i], nary[j])) return false;
        return true;
      default: {  // Disjunction::propagate (disjunction.rs:96-116)
        SKleene k[2] = {sub_eqyz(s, o), sub_eqyz(s, o + 3)};
        size_t num_disentailed = 0, unknown_formula = 0;
        for (size_t i = 0; i < 2; ++i) {
          if (k[i] == True) return true;
          if (k[i] == False) ++num_disentailed; else unknown_formula = i;
        }
        if (num_disentailed == 1) return prop_eqyz(s, o + 3 * unknown_formula);
        return num_disentailed != 2;
      }
    }
  }
  std::vector<Dep> dependencies() const override {
    std::vector<Dep> d;
    switch (kind) {
      case LessY: dep(d, o[0], Bound); dep(d, o[1], Bound); break;
      case NeqY: case EqY: dep(d, o[0], Inner); dep(d, o[1], Inner); break;
      case GreaterYPlusZ: case LessYPlusZ: case EqYPlusZ: case EqYMulZ:
        for (int i = 0; i < 3; ++i) dep(d, o[i], Bound);
        break;
      case DistinctN: case AllEqualN: for (auto& a : nary) dep(d, a, Inner); break;
      default:
        for (int i = 0; i < 6; ++i) dep(d, o[i], Bound);
        d = sorted_dedup(std::move(d));
    }
    return d;
  }
  Formula bclone() const override { return std::make_unique<FlatProp>(*this); }
};

inline Formula make_flat(int kind, const FOp* ops, int n) {
  auto p = std::make_unique<FlatProp>();
  p->kind = kind;
  if (kind == FlatProp::DistinctN || kind == FlatProp::AllEqualN) p->nary.assign(ops, ops + n);
  else for (int i = 0; i < n && i < 6; ++i) p->o[i] = ops[i];
  return p;
}

// ---------------------------------------------------------------------------
// propagation/store.rs -- the constraint store and THE fixpoint loop.
//   variant Faithful: reactor rebuilt at every consistency() incl. the
//     duplicate-subscription scan (store.rs:125-149, indexed_deps.rs:69-77).
//   variant Tuned: identical semantics/results; dependencies cached per
//     propagator, a static event-indexed CSR built once per model and filtered
//     by `active` at react time, no duplicate scan.  (Scheduling order after
//     an unsubscribe may differ from swap_remove order; results do not.)
// ---------------------------------------------------------------------------
enum class Variant { Faithful = 0, Tuned = 1 };

struct CStore {
  std::vector<Formula> propagators;
  std::vector<uint8_t> active;  // BitSet
  IndexedDeps reactor;
  RelaxedFifo scheduler;
  Variant variant = Variant::Faithful;
  uint64_t num_propagations = 0;  // added counter: propagate_one executions (store.rs:166)

  // Tuned-variant state
  std::vector<std::vector<Dep>> dep_cache;
  std::vector<std::vector<uint32_t>> static_deps;  // [kNumEvents*v + e] -> props (ascending)
  size_t static_props = 0;
  size_t live_subscriptions = 0;

  size_t size() const { return propagators.size(); }
  size_t alloc(Formula p) {  // store.rs:223-230
    propagators.push_back(std::move(p));
    active.push_back(1);
    return propagators.size() - 1;
  }

  // ---- Faithful ------------------------------------------------------------
  void init_reactor(const VStore& vstore) {  // store.rs:130-142
    reactor = IndexedDeps(vstore.size(), kNumEvents);
    for (size_t p = 0; p < propagators.size(); ++p) {
      if (!active[p]) continue;
      for (auto& d : propagators[p]->dependencies()) reactor.subscribe(d.first, d.second, p);
    }
  }
  void init_scheduler() {  // store.rs:144-149
    scheduler = RelaxedFifo(propagators.size());
    for (size_t p = 0; p < propagators.size(); ++p) if (active[p]) scheduler.schedule(p);
  }
  void unlink_prop(size_t p) {  // store.rs:200-207
    active[p] = 0;
    scheduler.unschedule(p);
    for (auto& d : propagators[p]->dependencies()) reactor.unsubscribe(d.first, d.second, p);
  }
  void react(VStore& vstore) {  // store.rs:191-198
    for (auto& d : vstore.drain_delta())
      for (size_t q : reactor.react(d.first, d.second)) scheduler.schedule(q);
  }
  SKleene propagator_consistency(size_t p, VStore& vstore) {  // store.rs:177-183
    if (propagators[p]->propagate(vstore)) return propagators[p]->is_subsumed(vstore);
    return False;
  }
  bool propagate_one(size_t p, VStore& vstore) {  // store.rs:166-175
    ++num_propagations;
    vstore.reset_changed();
    SKleene k = propagator_consistency(p, vstore);
    if (k == False) return false;
    if (k == True) { if (variant == Variant::Faithful) unlink_prop(p); else unlink_tuned(p); }
    else if (vstore.has_changed()) scheduler.schedule(p);  // store.rs:185-189
    return true;
  }
  bool propagation_loop(VStore& vstore) {  // store.rs:151-164
    bool consistent = true;
    while (!scheduler.is_empty() && consistent) {
      size_t p;
      while (scheduler.pop(&p)) {
        if (!propagate_one(p, vstore)) { consistent = false; break; }
        if (variant == Variant::Faithful) react(vstore); else react_tuned(vstore);
      }
    }
    return consistent;
  }

  // ---- Tuned ---------------------------------------------------------------
  void cache_deps_upto(size_t n) {
    while (dep_cache.size() < n) dep_cache.push_back(propagators[dep_cache.size()]->dependencies());
  }
  void prepare_tuned(const VStore& vstore) {
    PCPO_ASSERT(static_props <= propagators.size(), "tuned CSR ahead of the store (missing on_truncate)");
    cache_deps_upto(propagators.size());
    if (static_deps.size() < vstore.size() * kNumEvents) static_deps.resize(vstore.size() * kNumEvents);
    for (; static_props < propagators.size(); ++static_props)  // ids only grow => lists stay ascending
      for (auto& d : dep_cache[static_props]) {
        PCPO_ASSERT(d.first < vstore.size(), "dependency on a variable not in the vstore");
        static_deps[kNumEvents * d.first + d.second].push_back(uint32_t(static_props));
      }
    live_subscriptions = 0;
    for (size_t p = 0; p < propagators.size(); ++p) if (active[p]) live_subscriptions += dep_cache[p].size();
    scheduler = RelaxedFifo(propagators.size());
    for (size_t p = 0; p < propagators.size(); ++p) if (active[p]) scheduler.schedule(p);
  }
  void unlink_tuned(size_t p) {
    active[p] = 0;
    if (scheduler.inside_queue[p]) scheduler.unschedule(p);
    live_subscriptions -= dep_cache[p].size();
  }
  void react_tuned(VStore& vstore) {
    for (auto& d : vstore.drain_delta())
      for (size_t e = d.second; e < size_t(kNumEvents); ++e)
        for (uint32_t q : static_deps[kNumEvents * d.first + e])
          if (active[q]) scheduler.schedule(q);
  }
  // Label truncation (store.rs:320): the truncated ids are the largest ones, so
  // they sit at the back of their (ascending) CSR lists.
  void on_truncate(size_t n) {
    for (size_t p = static_props; p-- > n;)
      for (auto& d : dep_cache[p]) {
        auto& l = static_deps[kNumEvents * d.first + d.second];
        PCPO_ASSERT(!l.empty() && l.back() == p, "tuned CSR out of order");
        l.pop_back();
      }
    static_props = std::min(static_props, n);
    if (dep_cache.size() > n) dep_cache.resize(n);
  }

  // Subsumption for Store (store.rs:231-237): Kleene-and over *all* propagators
  SKleene is_subsumed(const VStore& vstore) const {
    SKleene k = True;
    for (auto& p : propagators) k = k_and(k, p->is_subsumed(vstore));
    return k;
  }
  // store.rs:247-257
  SKleene consistency(VStore& vstore) {
    bool consistent;
    bool all_unlinked;
    if (variant == Variant::Faithful) {
      init_reactor(vstore);
      init_scheduler();
      consistent = propagation_loop(vstore);
      all_unlinked = reactor.is_empty();
    } else {
      prepare_tuned(vstore);
      consistent = propagation_loop(vstore);
      all_unlinked = live_subscriptions == 0;
    }
    // The reference leaves undrained deltas behind on failure; they are
    // dropped with the failed space.  Drain so the next call starts clean.
    if (!consistent) { vstore.drain_delta(); return False; }
    return all_unlinked ? True : Unknown;
  }
};

// ---------------------------------------------------------------------------
// search/space.rs:21-66 + recomputation/no_recomputation.rs:28-62, with
// CopyMemory labels (copy_memory.rs:141-151) and cstore labels
// (propagation/store.rs:312-323: (len, active.clone())).
// ---------------------------------------------------------------------------
struct Label {
  std::shared_ptr<const std::vector<Dom>> domains;
  size_t num_props = 0;
  std::shared_ptr<const std::vector<uint8_t>> active;
};

struct Space {
  VStore vstore;
  CStore cstore;
  SKleene consistency() { return cstore.consistency(vstore); }  // space.rs:41-43
  Label label() const {
    Label l;
    l.domains = std::make_shared<const std::vector<Dom>>(vstore.memory);
    l.num_props = cstore.propagators.size();
    l.active = std::make_shared<const std::vector<uint8_t>>(cstore.active);
    return l;
  }
  void restore(const Label& l) {
    vstore.memory = *l.domains;
    vstore.delta_ev.assign(vstore.memory.size(), -1);
    vstore.delta_keys.clear();
    vstore.changed = false;
    cstore.propagators.resize(l.num_props);  // truncate (store.rs:320)
    cstore.active = *l.active;               // store.rs:321
    cstore.on_truncate(l.num_props);
  }
};

// ---------------------------------------------------------------------------
// search/ -- OneSolution o [Monitor o StopNode o BranchAndBound o] Propagation o
// Brancher(FirstSmallestVar, MiddleVal, BinarySplit), written as one explicit
// DFS loop with the same node order (one_solution.rs:46-51,92-105).
// ---------------------------------------------------------------------------
enum class NodeStatus : int { Satisfiable = 1, Unsatisfiable = -1, Unknown = 0, EndOfSearch = 2 };

// first_smallest_var.rs:30-39: first index among minimal size > 1.
inline long first_smallest_var(const VStore& vs) {
  long best = -1; uint32_t best_sz = 0;
  for (size_t i = 0; i < vs.size(); ++i) {
    uint32_t sz = vs.memory[i].size();
    if (sz > 1 && (best < 0 || sz < best_sz)) { best = long(i); best_sz = sz; }
  }
  return best;
}
// input_order.rs: first non-assigned variable.
inline long input_order_var(const VStore& vs) {
  for (size_t i = 0; i < vs.size(); ++i) if (vs.memory[i].size() > 1) return long(i);
  return -1;
}
// middle_val.rs:25-27 (truncating division).
inline int32_t middle_val(const Dom& d) { return (d.lower() + d.upper()) / 2; }
// min_val.rs
inline int32_t min_val(const Dom& d) { return d.lower(); }

struct Branch {
  Label label;
  int kind;       // 0: x <= v (binary_split.rs:46-51); 1: x > v (:52-57);
                  // 2: x == v, 3: x != v (enumerate.rs)
  size_t var; int32_t val;
};

struct Statistics {  // statistics.rs:19-53
  uint64_t num_solution = 0, num_failed_node = 0, num_prune = 0, num_nodes = 0;
};

enum class BBMode { None = 0, Minimize = 1, Maximize = 2 };

struct SearchConfig {
  uint64_t node_limit = 0;        // StopNode limit (0 = none), stop_node.rs:54-61
  bool all_solutions = false;     // AllSolution wrapper (all_solution.rs:41-50)
  int var_sel = 0;                // 0 FirstSmallestVar, 1 InputOrder
  int val_sel = 0;                // 0 MiddleVal, 1 MinVal
  int distributor = 0;            // 0 BinarySplit, 1 Enumerate
  BBMode bb_mode = BBMode::None;  // branch_and_bound.rs:69-94
  size_t bb_var = 0;
};

struct NodeRecord { int status; uint64_t dom_hash; };

inline uint64_t hash_domains(const std::vector<Dom>& m) {  // FNV-1a over the (lb, ub) words of every maximal run
  uint64_t h = 1469598103934665603ull;
  for (auto& d : m)
    d.for_each_run([&](int32_t lb, int32_t ub) {
      for (uint32_t w : {uint32_t(lb), uint32_t(ub)})
        for (int b = 0; b < 4; ++b) { h ^= (w >> (8 * b)) & 0xffu; h *= 1099511628211ull; }
    });
  return h;
}

struct Search {
  SearchConfig cfg;
  Statistics stats;
  std::vector<Branch> queue;  // VectorStack
  bool started = false;
  bool has_bb_value = false; int32_t bb_value = 0;
  uint64_t nodes_explored = 0;  // StopNode counter
  std::function<void(const Space&, int status)> on_node;  // trace hook (called after consistency)

  bool flat = false;  // allocate FlatProp descriptors instead of boxed view trees

  void apply_alternative(Space& sp, const Branch& b) {
    if (flat) {
      FOp x{int32_t(b.var), 0}, c{-1, b.val}, c1{-1, b.val + 1};
      FOp leq[2] = {x, c1}, gt[2] = {c, x}, xc[2] = {x, c};
      switch (b.kind) {
        case 0: sp.cstore.alloc(make_flat(FlatProp::LessY, leq, 2)); break;
        case 1: sp.cstore.alloc(make_flat(FlatProp::LessY, gt, 2)); break;
        case 2: sp.cstore.alloc(make_flat(FlatProp::EqY, xc, 2)); break;
        default: sp.cstore.alloc(make_flat(FlatProp::NeqY, xc, 2)); break;
      }
      return;
    }
    auto x = [&] { return Var(std::make_unique<Identity>(b.var)); };
    auto v = [&] { return Var(std::make_unique<Constant>(b.val)); };
    switch (b.kind) {
      case 0: sp.cstore.alloc(make_x_leq_y(x(), v())); break;
      case 1: sp.cstore.alloc(make_x_greater_y(x(), v())); break;
      case 2: sp.cstore.alloc(std::make_unique<XEqY>(x(), v())); break;
      default: sp.cstore.alloc(std::make_unique<XNeqY>(x(), v())); break;
    }
  }

  // Propagation::enter + Brancher::enter (+ BranchAndBound, StopNode, Monitor).
  NodeStatus enter_child(Space& sp) {
    if (cfg.bb_mode != BBMode::None && has_bb_value && flat) {
      FOp v{int32_t(cfg.bb_var), 0}, b{-1, bb_value};
      FOp mn[2] = {v, b}, mx[2] = {b, v};
      sp.cstore.alloc(make_flat(FlatProp::LessY, cfg.bb_mode == BBMode::Minimize ? mn : mx, 2));
    } else if (cfg.bb_mode != BBMode::None && has_bb_value) {  // branch_and_bound.rs:76-87
      Var var = std::make_unique<Identity>(cfg.bb_var);
      Var bound = std::make_unique<Constant>(bb_value);
      if (cfg.bb_mode == BBMode::Minimize) sp.cstore.alloc(std::make_unique<XLessY>(std::move(var), std::move(bound)));
      else sp.cstore.alloc(make_x_greater_y(std::move(var), std::move(bound)));
    }
    SKleene k = sp.consistency();  // propagation.rs:49
    NodeStatus st;
    std::vector<Branch> branches;
    if (k == True) st = NodeStatus::Satisfiable;
    else if (k == False) st = NodeStatus::Unsatisfiable;
    else {
      st = NodeStatus::Unknown;
      long var = cfg.var_sel == 0 ? first_smallest_var(sp.vstore) : input_order_var(sp.vstore);
      PCPO_ASSERT(var >= 0, "Cannot select a variable in a space where all variables are assigned.");
      const Dom& dom = sp.vstore[size_t(var)];
      int32_t val = cfg.val_sel == 0 ? middle_val(dom) : min_val(dom);
      Label l = sp.label();  // Branch::distribute: freeze + label per alternative
      int k0 = cfg.distributor == 0 ? 0 : 2;
      branches.push_back(Branch{l, k0, size_t(var), val});
      branches.push_back(Branch{l, k0 + 1, size_t(var), val});
    }
    if (on_node) on_node(sp, int(k));
    if (st == NodeStatus::Satisfiable && cfg.bb_mode != BBMode::None) {  // branch_and_bound.rs:89-92
      has_bb_value = true; bb_value = sp.vstore[cfg.bb_var].lower();
    }
    ++nodes_explored;  // stop_node.rs:54-61
    bool stop = cfg.node_limit && nodes_explored >= cfg.node_limit;
    // Monitor::enter (monitor.rs:61-66) sees the status after StopNode.
    ++stats.num_nodes;
    if (stop) st = NodeStatus::EndOfSearch;
    else if (st == NodeStatus::Satisfiable) ++stats.num_solution;
    else if (st == NodeStatus::Unsatisfiable) ++stats.num_failed_node;
    if (st == NodeStatus::Unknown)
      for (auto it = branches.rbegin(); it != branches.rend(); ++it) queue.push_back(std::move(*it));
    return st;
  }

  // OneSolution::enter (one_solution.rs:92-105).
  NodeStatus one_solution(Space& sp) {
    if (queue.empty() && started) return NodeStatus::EndOfSearch;
    NodeStatus status = NodeStatus::Unsatisfiable;
    auto fold = [&](NodeStatus child) {  // enter_child's match, one_solution.rs:56-64
      if (child == NodeStatus::Satisfiable || child == NodeStatus::EndOfSearch) status = child;
    };
    if (queue.empty() && !started) { started = true; fold(enter_child(sp)); }
    while (status != NodeStatus::EndOfSearch && status != NodeStatus::Satisfiable && !queue.empty()) {
      Branch b = std::move(queue.back());
      queue.pop_back();
      sp.restore(b.label);  // Branch::commit (branch.rs:51-55)
      apply_alternative(sp, b);
      fold(enter_child(sp));
    }
    return status;
  }

  NodeStatus run(Space& sp) {
    if (!cfg.all_solutions) return one_solution(sp);
    NodeStatus st = one_solution(sp);  // all_solution.rs:41-50
    while (st != NodeStatus::EndOfSearch) st = one_solution(sp);
    return st;
  }
};

