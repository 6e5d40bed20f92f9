// pcp_eval.cuh -- per-propagator evaluation (propagate + is_subsumed) on the device.
// Included by pcp_device.cuh after Params / the memory helpers.
//
// Every propagator is evaluated in two steps: a *pure* step that reads the operand views
// and computes the narrowed bounds exactly as the reference's `propagate` would (file:line
// at each function), staging at most one update per operand; and `apply_updates`, which
// issues the staged updates as fire-and-forget reductions on the bounds and on the dirty
// bit set.  A cheap inline "would this evaluation change anything?" test keeps the common
// case of a sweep free of calls.
// (No include guard and no namespace of its own: pcp_body.cuh includes this file inside the
// namespace of the kernel variant being compiled.)


struct IV { int lo, hi; };

// Per-thread context of one fixpoint launch (read-mostly; counters live in registers).
struct Ctx {
  const Params* P;
  int2* sdom;             // shared-memory snapshot (or nullptr)
  uint32_t sdom_s;        // its shared-space address (hot reads are explicit LDS, not generic loads)
  uint32_t* next_bits;    // dirty bit set written in this iteration (read in the next one)
  bool local;             // updates go to the shared-memory snapshot only (posted props, per CTA)
  bool mark_dirty;        // narrowed variables enter the worklist
  bool bookkeep;          // entailed propagators are deactivated + trailed
  bool mirror;            // row-local rounds: every update is also applied to this CTA's snapshot
  int* flags;             // shared: [0] this CTA narrowed a variable, [1] saw a failure
  // what the update paths of the lean sweeps need (CTA-uniform, so kept here and not in the
  // by-value FamSweep: past 128 bytes that struct travels through local memory)
  int2* dom_w;            // the store itself
  uint32_t* trail;        // entailment trail and its length
  unsigned* trail_cnt;
  uint32_t flags_s;       // shared-space addresses: flags, the sweep's TrailBuf, its dirty bitmap (0: none)
  uint32_t tbuf_s;
  uint32_t dbm_s;
};

__device__ __forceinline__ void set_failed(const Ctx& c) { c.flags[1] = 1; }

// warp-aggregated append to the entailment trail + clear of the active bit
// (Store::unlink_prop, propagation/store.rs:200-207)
__device__ __forceinline__ void deactivate(const Ctx& c, uint32_t* active, unsigned fam, int slot) {
  unsigned bit = 1u << (slot & 31);
  unsigned old = atomicAnd(&active[slot >> 5], ~bit);
  if (!(old & bit)) return;
  unsigned m = __activemask();
  int leader = __ffs(m) - 1;
  int lane = threadIdx.x & 31;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(&c.P->ctl->trail_cnt, (unsigned)__popc(m));
  base = __shfl_sync(m, base, leader);
  c.P->trail[base + __popc(m & lanemask_lt())] = make_ref(fam, (unsigned)slot);
}

// Domain readers.  `SMEM` reads the CTA's snapshot, otherwise L2 (ld.global.cg).
// The snapshot is accessed through explicit shared-space instructions: through the generic
// pointer in Ctx the compiler cannot prove the address space and emits generic loads.
__device__ __forceinline__ int2 lds_dom(uint32_t addr) {
  int2 r;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_dom(uint32_t addr, int lo, int hi) {
  asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi));
}
template <bool SMEM>
__device__ __forceinline__ int2 rd_var(const Ctx& c, int var) {
  return SMEM ? lds_dom(c.sdom_s + 8u * (unsigned)var) : ldcg_dom(&c.P->dom[var]);
}
template <bool SMEM>
__device__ __noinline__ IV rd_sum(const Ctx& c, int sum_id, int off) {
  // Sum::read (term/sum.rs:71-82): [sum lo_i, sum hi_i] over the terms.  A sum with more than
  // one term is never written: its update is an overlap test (term/sum.rs:62-69), which the
  // staging step performs for every read-only view (var < 0).
  const Params& P = *c.P;
  const int b = __ldg(&P.sum_ptr[sum_id]), e = __ldg(&P.sum_ptr[sum_id + 1]);
  int lo = off, hi = off;
  for (int t = b; t < e; ++t) {
    int2 term = __ldg(&P.sum_terms[t]);
    if (term.x >= 0) {
      int2 d = rd_var<SMEM>(c, term.x);
      lo += d.x + term.y;
      hi += d.y + term.y;
    } else {
      lo += term.y;
      hi += term.y;
    }
  }
  return IV{lo, hi};
}
template <bool SMEM>
__device__ __forceinline__ IV rd(const Ctx& c, int var, int off) {
  if (var < 0) {
    if (var == -1) return IV{off, off};  // Constant (term/constant.rs:55-63)
    return rd_sum<SMEM>(c, -2 - var, off);
  }
  int2 d = rd_var<SMEM>(c, var);
  return IV{d.x + off, d.y + off};   // Addition (term/addition.rs:93-101)
}

#ifdef PCP_SET
// ---------------------------------------------------------------------------------------
// IntervalSet domains: the bit set of a variable (one bit per value of the window
// [bits_base, bits_base + 32 * bits_W)) lives in HBM / L2 beside the cached bounds; the domain
// is bits /\ [lo, hi].  Bits are only ever cleared (RED.AND / ATOM.AND), bounds only ever
// tightened, so every view a thread can form from (possibly stale) reads is a superset of the
// current domain -- exactly the situation the parity argument of the bounds already covers.
// `raw` values are in variable space (view value minus the view's offset).
// ---------------------------------------------------------------------------------------
struct SetW { uint32_t* bits; int W; int base; };
__device__ __forceinline__ SetW set_of(const Params& P) { return SetW{P.bits, P.bits_W, P.bits_base}; }
__device__ __forceinline__ uint32_t* sb_words(const SetW& s, int var) { return s.bits + (size_t)var * (unsigned)s.W; }
__device__ __forceinline__ bool sb_test(const SetW& s, int var, int raw) {
  const unsigned b = (unsigned)(raw - s.base);
  return (__ldcg(sb_words(s, var) + (b >> 5)) >> (b & 31)) & 1u;
}
// smallest element in [from, hi], INT32_MAX if there is none (IntervalSet::shrink_left)
__device__ __noinline__ int sb_next(SetW s, int var, int from, int hi) {
  if (from > hi) return INT32_MAX;
  const uint32_t* w = sb_words(s, var);
  const unsigned b = (unsigned)(from - s.base), e = (unsigned)(hi - s.base);
  unsigned i = b >> 5;
  unsigned word = __ldcg(w + i) & (0xffffffffu << (b & 31));
  while (true) {
    if (word) { const unsigned pos = i * 32u + (unsigned)(__ffs(word) - 1); return pos <= e ? (int)pos + s.base : INT32_MAX; }
    if (++i > (e >> 5)) return INT32_MAX;
    word = __ldcg(w + i);
  }
}
// largest element in [lo, from], INT32_MIN if there is none (IntervalSet::shrink_right)
__device__ __noinline__ int sb_prev(SetW s, int var, int from, int lo) {
  if (from < lo) return INT32_MIN;
  const uint32_t* w = sb_words(s, var);
  const unsigned b = (unsigned)(from - s.base), e = (unsigned)(lo - s.base);
  int i = (int)(b >> 5);
  unsigned word = __ldcg(w + i) & (0xffffffffu >> (31u - (b & 31)));
  while (true) {
    if (word) { const unsigned pos = (unsigned)i * 32u + 31u - (unsigned)__clz(word); return pos >= e ? (int)pos + s.base : INT32_MIN; }
    if (--i < (int)(e >> 5)) return INT32_MIN;
    word = __ldcg(w + i);
  }
}
// IntervalSet::difference(&v) for a value strictly inside the bounds: true when it was there
__device__ __forceinline__ bool sb_clear(const SetW& s, int var, int raw) {
  const unsigned b = (unsigned)(raw - s.base);
  return (atomicAnd(sb_words(s, var) + (b >> 5), ~(1u << (b & 31))) >> (b & 31)) & 1u;
}
// the 32 bits for raw values [raw, raw + 32)  (the word behind the last one of a row is the next
// row or the array's padding: the callers mask it away)
__device__ __forceinline__ unsigned sb_get32(const SetW& s, int var, int raw) {
  const unsigned b = (unsigned)(raw - s.base);
  const uint32_t* w = sb_words(s, var) + (b >> 5);
  return __funnelshift_r(__ldcg(w), __ldcg(w + 1), b & 31);
}
// Cardinality::size()
__device__ __forceinline__ unsigned sb_count(const SetW& s, int var, int lo, int hi) {
  unsigned n = 0;
  for (long long r = lo; r <= hi; r += 32) {
    const int left = (int)min(32ll, (long long)hi - r + 1);
    n += __popc(sb_get32(s, var, (int)r) & (left == 32 ? 0xffffffffu : ((1u << left) - 1u)));
  }
  return n;
}
// Are the views  bits[xv] /\ [xlo, xhi] + xo  and  bits[yv] /\ [ylo, yhi] + yo  disjoint as sets
// (IntervalSet::is_disjoint, what XEqY::is_subsumed asks at x_eq_y.rs:89)?  Bounds are raw.
__device__ __noinline__ bool sb_disjoint(SetW s, int xv, int xlo, int xhi, int xo, int yv, int ylo, int yhi, int yo) {
  const int L = max(xlo + xo, ylo + yo), U = min(xhi + xo, yhi + yo);
  for (long long t = L; t <= U; t += 32) {
    const int left = (int)min(32ll, (long long)U - t + 1);
    const unsigned m = left == 32 ? 0xffffffffu : ((1u << left) - 1u);
    if (sb_get32(s, xv, (int)t - xo) & sb_get32(s, yv, (int)t - yo) & m) return false;
  }
  return true;
}
// Cheap proof that two overlapping non-singleton views share a value: the larger of the two
// lower bounds belongs to its own set (the bounds are kept on elements), so one bit test in the
// other set settles the common case of dense domains; the smaller upper bound is the second
// try.  false = no witness found (not a proof of disjointness).
__device__ __forceinline__ bool sb_witness(const SetW& s, int xv, int xlo, int xhi, int xo, int yv, int ylo, int yhi, int yo) {
  const int L = max(xlo + xo, ylo + yo), U = min(xhi + xo, yhi + yo);
  const bool lx = xlo + xo >= ylo + yo;           // L is x's lower bound
  if (sb_test(s, lx ? yv : xv, lx ? L - yo : L - xo)) return true;
  const bool ux = xhi + xo <= yhi + yo;           // U is x's upper bound
  return sb_test(s, ux ? yv : xv, ux ? U - yo : U - xo);
}
#endif  // PCP_SET

// ---------------------------------------------------------------------------------------
// staged updates: variable/store.rs:151-166 through term/addition.rs:80-90 /
// term/constant.rs:43-53.  All bounds are in view space.
// ---------------------------------------------------------------------------------------
struct Upd { int var, off, cur_lo, cur_hi, nlo, nhi; };
struct UpdSet {   // slot 0 = x, 1 = y, 2 = z: constant indices only, so the set lives in registers
  Upd u[3];
  bool on[3];
#ifdef PCP_SET
  int rm_var, rm_raw;   // IntervalSet::difference of a value inside the bounds (rm_var < 0: none)
#endif
};
__device__ __forceinline__ void upd_clear(UpdSet& us) {
  us.on[0] = us.on[1] = us.on[2] = false;
#ifdef PCP_SET
  us.rm_var = -1;
#endif
}

// Stage "view <- [nlo, nhi]" (already intersected with `cur`).  Returns false when the new
// domain is empty (the reference's update -> false); a Constant is never written but fails
// when it falls outside (nlo > nhi covers it because cur is the singleton).
template <int I>
__device__ __forceinline__ bool stage(const Ctx& c, UpdSet& us, int var, int off, IV cur, int nlo, int nhi) {
  if (nlo > nhi) return false;
  if (var >= 0 && (nlo > cur.lo || nhi < cur.hi)) {
#ifdef PCP_SET
    // IntervalSet::shrink_left / shrink_right: a bound that moves lands on the next value that
    // is still in the set; nothing left between the new bounds = the empty domain
    const SetW sw = set_of(*c.P);
    if (nlo > cur.lo) { const int r = sb_next(sw, var, nlo - off, nhi - off); if (r == INT32_MAX) return false; nlo = r + off; }
    if (nhi < cur.hi) { const int r = sb_prev(sw, var, nhi - off, nlo - off); if (r == INT32_MIN) return false; nhi = r + off; }
#endif
    Upd& r = us.u[I];
    us.on[I] = true;
    r.var = var; r.off = off; r.cur_lo = cur.lo; r.cur_hi = cur.hi; r.nlo = nlo; r.nhi = nhi;
  }
  return true;
}

// All updates are fire-and-forget reductions (RED.MAX / RED.MIN on the bounds, RED.OR on the
// dirty bit set): nothing in the evaluation waits for an L2 round trip.  A variable is marked
// dirty whenever this thread's view says the update narrows it; if a concurrent update had
// already narrowed it further, that thread marked it too.  A domain emptied by two concurrent
// updates (lo from one thread, hi from another) is caught when the dirty variables are
// re-read at the start of the next iteration -- there always is one, flags[0] is set.
__device__ __forceinline__ void apply_updates(const Ctx& c, const UpdSet& us) {
#ifdef PCP_SET
  if (us.rm_var >= 0 && !c.local) {  // (a CTA-local evaluation leaves the shared bit set alone)
    if (sb_clear(set_of(*c.P), us.rm_var, us.rm_raw) && c.mark_dirty) {
      atomicOr(&c.next_bits[us.rm_var >> 5], 1u << (us.rm_var & 31));
      c.flags[0] = 1;
    }
  }
#endif
  if (!(us.on[0] || us.on[1] || us.on[2])) return;
  const Params& P = *c.P;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (!us.on[i]) continue;
    const Upd& r = us.u[i];
    const int nl = r.nlo - r.off, nh = r.nhi - r.off;
    if (c.local) { sts_dom(c.sdom_s + 8u * (unsigned)r.var, nl, nh); continue; }
    int2* d = &P.dom[r.var];
    if (r.nlo > r.cur_lo) atomicMax(&d->x, nl);
    if (r.nhi < r.cur_hi) atomicMin(&d->y, nh);
    if (c.mirror) {  // keep the CTA's snapshot current for the next local round
      if (r.nlo > r.cur_lo) atomicMax(&c.sdom[r.var].x, nl);
      if (r.nhi < r.cur_hi) atomicMin(&c.sdom[r.var].y, nh);
    }
    if (c.mark_dirty) atomicOr(&c.next_bits[r.var >> 5], 1u << (r.var & 31));
  }
  if (!c.local && c.mark_dirty) c.flags[0] = 1;
}

// Single update (n-ary propagators write their operands one by one).
__device__ __forceinline__ bool tighten(const Ctx& c, int var, int off, IV cur, int nlo, int nhi) {
  UpdSet us;
  upd_clear(us);
  if (!stage<0>(c, us, var, off, cur, nlo, nhi)) return false;
  apply_updates(c, us);
  return true;
}

// Result of evaluating one propagator: propagate + is_subsumed (store.rs:177-183).
enum Eval : int { E_FAIL = -1, E_UNKNOWN = 0, E_ENTAILED = 1 };

// the x operand shares descriptor word 0 with the kind: 28 bits, the top of the range encodes
// Constant (kConstVar28) and sum views (kSumBase28 + sum id)
__device__ __forceinline__ int dec_var28(unsigned w0) {
  unsigned v = w0 & kConstVar28;
  if (v < kSumBase28) return (int)v;
  return v == kConstVar28 ? -1 : -2 - (int)(v - kSumBase28);
}

// --- binary family: XLessY / XNeqY / XEqY ---------------------------------------------------
// `x`, `y`: the operand views as read by the caller (each propagate reads its views once).
#ifdef PCP_SET
// XEqY::propagate on IntervalSet (x_eq_y.rs:102-107: both sides <- x /\ y as sets), evaluated
// directly on the shared bit sets: the values of one side that the other side lacks are
// cleared, the bounds move to the extreme common values.  Is_subsumed on the result
// (x_eq_y.rs:84-93): one common value left -> True, none -> False.
__device__ __noinline__ Eval eq_sets(const Ctx& c, int xv, int xo, IV x, int yv, int yo, IV y, UpdSet& us) {
  const SetW sw = set_of(*c.P);
  const int L = max(x.lo, y.lo), U = min(x.hi, y.hi);
  int first = INT32_MAX, last = INT32_MIN;
  bool chx = false, chy = false;
  for (long long t = L; t <= U; t += 32) {
    const int left = (int)min(32ll, (long long)U - t + 1);
    const unsigned m = left == 32 ? 0xffffffffu : ((1u << left) - 1u);
    const unsigned a = sb_get32(sw, xv, (int)t - xo) & m, b = sb_get32(sw, yv, (int)t - yo) & m;
    const unsigned both = a & b;
    if (both) { first = min(first, (int)t + __ffs(both) - 1); last = (int)t + 31 - __clz(both); }
    if (!c.local) {
      for (unsigned r = a & ~both; r; r &= r - 1) chx |= sb_clear(sw, xv, (int)t - xo + __ffs(r) - 1);
      for (unsigned r = b & ~both; r; r &= r - 1) chy |= sb_clear(sw, yv, (int)t - yo + __ffs(r) - 1);
    }
  }
  if (first == INT32_MAX) return E_FAIL;
  if (c.mark_dirty && !c.local) {
    if (chx) atomicOr(&c.next_bits[xv >> 5], 1u << (xv & 31));
    if (chy) atomicOr(&c.next_bits[yv >> 5], 1u << (yv & 31));
    if (chx || chy) c.flags[0] = 1;
  }
  if (!stage<0>(c, us, xv, xo, x, first, last)) return E_FAIL;
  if (!stage<1>(c, us, yv, yo, y, first, last)) return E_FAIL;
  return first == last ? E_ENTAILED : E_UNKNOWN;
}
// XNeqY on IntervalSet (x_neq_y.rs:82-93 with IntervalSet::difference, is_subsumed = not XEqY,
// x_neq_y.rs:71-73 / x_eq_y.rs:84-93 with the set is_disjoint).  `s` is the assigned side
// (value v in view space), `t` the other one.
__device__ __forceinline__ Eval neq_assigned(const Ctx& c, UpdSet& us, int v, int tv, int to, IV t, bool t_is_x) {
  if (tv < 0) return t.lo == v ? E_FAIL : E_ENTAILED;         // Constant: nothing to narrow
  if (v < t.lo || v > t.hi) return E_ENTAILED;                // not in the other domain: disjoint
  if (t.lo == t.hi) return E_FAIL;                            // two equal singletons
  if (v == t.lo || v == t.hi) {                               // a bound goes: Bound / Assignment event
    const int nlo = v == t.lo ? t.lo + 1 : t.lo, nhi = v == t.hi ? t.hi - 1 : t.hi;
    const bool ok = t_is_x ? stage<0>(c, us, tv, to, t, nlo, nhi) : stage<1>(c, us, tv, to, t, nlo, nhi);
    return ok ? E_ENTAILED : E_FAIL;
  }
  us.rm_var = tv;                                             // an interior value goes: Inner event
  us.rm_raw = v - to;
  return E_ENTAILED;
}
__device__ __forceinline__ Eval eval_bin_set(const Ctx& c, unsigned kind, int xv, int xo, int yv, int yo, IV x, IV y, UpdSet& us) {
  if (kind == B_NEQ) {
    if (x.lo == x.hi) return neq_assigned(c, us, x.lo, yv, yo, y, false);
    if (y.lo == y.hi) return neq_assigned(c, us, y.lo, xv, xo, x, true);
    if (x.hi < y.lo || y.hi < x.lo) return E_ENTAILED;
    const SetW sw = set_of(*c.P);
    if (sb_witness(sw, xv, x.lo - xo, x.hi - xo, xo, yv, y.lo - yo, y.hi - yo, yo)) return E_UNKNOWN;
    return sb_disjoint(sw, xv, x.lo - xo, x.hi - xo, xo, yv, y.lo - yo, y.hi - yo, yo) ? E_ENTAILED : E_UNKNOWN;
  }
  // B_EQ
  if (x.hi < y.lo || y.hi < x.lo) return E_FAIL;
  if (xv < 0 || yv < 0) {  // against a Constant: the other side becomes that value if it has it
    const int v = xv < 0 ? x.lo : y.lo;
    if (xv < 0 && yv < 0) return E_ENTAILED;  // (equal, or the test above failed)
    const bool ok = xv < 0 ? stage<1>(c, us, yv, yo, y, v, v) : stage<0>(c, us, xv, xo, x, v, v);
    if (!ok) return E_FAIL;
    // stage only looks at the set when a bound moves: an assigned side is in its set by invariant
    return E_ENTAILED;
  }
  return eq_sets(c, xv, xo, x, yv, yo, y, us);
}
#endif

__device__ __forceinline__ Eval eval_bin(const Ctx& c, int4 d, IV x, IV y, UpdSet& us) {
  unsigned kind = (unsigned)d.x >> 28;
  int xv = dec_var28((unsigned)d.x), xo = d.y, yv = d.z, yo = d.w;
#ifdef PCP_SET
  if (kind != B_LESS) return eval_bin_set(c, kind, xv, xo, yv, yo, x, y, us);
#endif
  // A multi-term Sum operand is never narrowed (term/sum.rs:62-69): is_subsumed, which the
  // store evaluates after propagate (store.rs:177-183), re-reads it unchanged.
  const bool xro = xv <= -2, yro = yv <= -2;
  IV nx = x, ny = y;
  if (kind == B_NEQ) {  // cmp/x_neq_y.rs:82-93 + Interval::difference
    if (x.lo == x.hi) {
      if (ny.lo == x.lo) ny.lo++; else if (ny.hi == x.lo) ny.hi--;
    } else if (y.lo == y.hi) {
      if (nx.lo == y.lo) nx.lo++; else if (nx.hi == y.lo) nx.hi--;
    }
    if (!stage<1>(c, us, yv, yo, y, ny.lo, ny.hi)) return E_FAIL;
    if (!stage<0>(c, us, xv, xo, x, nx.lo, nx.hi)) return E_FAIL;
  } else if (kind == B_LESS) {  // cmp/x_less_y.rs:101-108
    nx.hi = min(x.hi, y.hi - 1);
    if (!stage<0>(c, us, xv, xo, x, nx.lo, nx.hi)) return E_FAIL;
    ny.lo = max(y.lo, x.lo + 1);
    if (!stage<1>(c, us, yv, yo, y, ny.lo, ny.hi)) return E_FAIL;
  } else {  // B_EQ: cmp/x_eq_y.rs:102-107
    nx.lo = ny.lo = max(x.lo, y.lo);
    nx.hi = ny.hi = min(x.hi, y.hi);
    if (!stage<0>(c, us, xv, xo, x, nx.lo, nx.hi)) return E_FAIL;
    if (!stage<1>(c, us, yv, yo, y, ny.lo, ny.hi)) return E_FAIL;
  }
  const IV px = xro ? x : nx, py = yro ? y : ny;  // what is_subsumed reads back
  if (kind == B_LESS) {  // x_less_y.rs:84-91
    if (px.lo >= py.hi) return E_FAIL;
    return px.hi < py.lo ? E_ENTAILED : E_UNKNOWN;
  }
  // XEqY::is_subsumed (x_eq_y.rs:84-93); XNeqY negates it (x_neq_y.rs:71-73)
  const bool same_singleton = px.lo == py.hi && px.hi == py.lo;
  const bool disjoint = px.hi < py.lo || py.hi < px.lo;
  if (kind == B_NEQ) return same_singleton ? E_FAIL : (disjoint ? E_ENTAILED : E_UNKNOWN);
  return same_singleton ? E_ENTAILED : (disjoint ? E_FAIL : E_UNKNOWN);
}
// true when evaluating the propagator would change nothing: no pruning, no failure,
// not entailed (the common case of a sweep; keeps the hot loop free of calls)
#ifdef PCP_SET
// IntervalSet: an XNeqY between two unassigned, overlapping views is a no-op only when they are
// known to share a value (else it may be entailed: decided exactly out of line); an XEqY can
// always have interior values to remove.
__device__ __forceinline__ bool bin_is_noop_set(const Ctx& c, int4 d, IV x, IV y) {
  const unsigned kind = (unsigned)d.x >> 28;
  if (kind == B_LESS) return y.hi > x.hi && x.lo < y.lo && x.hi >= y.lo;
  if (kind == B_EQ) return false;
  const int xv = dec_var28((unsigned)d.x), yv = d.z;
  if (xv < 0 || yv < 0 || x.lo == x.hi || y.lo == y.hi || x.hi < y.lo || y.hi < x.lo) return false;
  return sb_witness(set_of(*c.P), xv, x.lo - d.y, x.hi - d.y, d.y, yv, y.lo - d.w, y.hi - d.w, d.w);
}
#endif
__device__ __forceinline__ bool bin_is_noop(unsigned kind, IV x, IV y) {
  if (kind == B_NEQ) return x.lo != x.hi && y.lo != y.hi && !(x.hi < y.lo || y.hi < x.lo);
  if (kind == B_LESS) return y.hi > x.hi && x.lo < y.lo && x.hi >= y.lo;
  return x.lo == y.lo && x.hi == y.hi && x.lo < x.hi;
}

// --- ternary family ------------------------------------------------------------------------
struct Tri { int xv, xo, yv, yo, zv, zo; };

// XGreaterYPlusZ::propagate on local copies (x_greater_y_plus_z.rs:106-119); `strict`=1 for
// x > y+z, 0 for x >= y+z (the Addition(x,1) of cmp/mod.rs:73).  Pure: narrows x, y, z.
__device__ __forceinline__ bool prop_greater(IV& x, IV& y, IV& z, int strict) {
  int nxlo = max(x.lo, y.lo + z.lo + strict);
  int nyhi = min(y.hi, x.hi - z.lo - strict);
  int nzhi = min(z.hi, x.hi - y.lo - strict);
  if (nxlo > x.hi || y.lo > nyhi || z.lo > nzhi) return false;
  x.lo = nxlo; y.hi = nyhi; z.hi = nzhi;
  return true;
}
// XLessYPlusZ::propagate (x_less_y_plus_z.rs:106-120)
__device__ __forceinline__ bool prop_less(IV& x, IV& y, IV& z, int strict) {
  int nxhi = min(x.hi, y.hi + z.hi - strict);
  int nylo = max(y.lo, x.lo - z.hi + strict);
  int nzlo = max(z.lo, x.lo - y.hi + strict);
  if (x.lo > nxhi || nylo > y.hi || nzlo > z.hi) return false;
  x.hi = nxhi; y.lo = nylo; z.lo = nzlo;
  return true;
}
// Kleene entailment tests (x_greater_y_plus_z.rs:84-98, x_less_y_plus_z.rs:84-98,
// x_eq_y_plus_z.rs:56-58)
__device__ __forceinline__ int sub_greater(IV x, IV y, IV z, int strict) {
  if (x.hi < y.lo + z.lo + strict) return -1;
  if (x.lo >= y.hi + z.hi + strict) return 1;
  return 0;
}
__device__ __forceinline__ int sub_less(IV x, IV y, IV z, int strict) {
  if (x.lo > y.hi + z.hi - strict) return -1;
  if (x.hi <= y.lo + z.lo - strict) return 1;
  return 0;
}
__device__ __forceinline__ int sub_eq(IV x, IV y, IV z) {
  return min(sub_greater(x, y, z, 0), sub_less(x, y, z, 0));
}
// Operands that are multi-term Sum views are checked (overlap) but never narrowed, so whoever
// reads them next -- the second half of XEqYPlusZ, is_subsumed -- sees the original interval.
struct TriRo { bool x, y, z; };
__device__ __forceinline__ TriRo tri_ro(const Tri& t) { return TriRo{t.xv <= -2, t.yv <= -2, t.zv <= -2}; }
__device__ __forceinline__ void undo_ro(const TriRo& ro, IV& x, IV& y, IV& z, IV x0, IV y0, IV z0) {
  if (ro.x) x = x0;
  if (ro.y) y = y0;
  if (ro.z) z = z0;
}
// XEqYPlusZ::propagate = geq then leq re-reading the store (x_eq_y_plus_z.rs:79-81)
__device__ __forceinline__ bool prop_eq(IV& x, IV& y, IV& z, const TriRo& ro) {
  const IV x0 = x, y0 = y, z0 = z;
  if (!prop_greater(x, y, z, 0)) return false;
  undo_ro(ro, x, y, z, x0, y0, z0);
  if (!prop_less(x, y, z, 0)) return false;
  undo_ro(ro, x, y, z, x0, y0, z0);
  return true;
}
__device__ __forceinline__ bool stage_tri(const Ctx& c, UpdSet& us, const Tri& t, IV x0, IV y0, IV z0, IV x, IV y, IV z) {
  return stage<0>(c, us, t.xv, t.xo, x0, x.lo, x.hi) && stage<1>(c, us, t.yv, t.yo, y0, y.lo, y.hi) &&
         stage<2>(c, us, t.zv, t.zo, z0, z.lo, z.hi);
}

// XEqYMulZ (cmp/x_eq_y_mul_z.rs:68-116): the interval product y * z = [min, max] of the four
// bound products (64-bit: bounds stay below 2^29 in magnitude).
struct IV64 { long long lo, hi; };
__device__ __forceinline__ IV64 mul_iv(IV y, IV z) {
  const long long p0 = (long long)y.lo * z.lo, p1 = (long long)y.lo * z.hi, p2 = (long long)y.hi * z.lo,
                  p3 = (long long)y.hi * z.hi;
  return IV64{min(min(p0, p1), min(p2, p3)), max(max(p0, p1), max(p2, p3))};
}
__device__ __forceinline__ Eval eval_ter(const Ctx& c, int4 a, int2 b, IV x0, IV y0, IV z0, UpdSet& us) {
  unsigned kind = (unsigned)a.x >> 28;
  Tri t{dec_var28((unsigned)a.x), a.y, a.z, a.w, b.x, b.y};
  IV x = x0, y = y0, z = z0;
  const TriRo ro = tri_ro(t);
  int s;
  if (kind == T_MUL) {
    // propagate: x <- x /\ (y * z), y and z untouched (x_eq_y_mul_z.rs:99-105); is_subsumed on the
    // result: no overlap -> False; y * z and x both singletons -> True (x_eq_y_mul_z.rs:73-90)
    const IV64 yz = mul_iv(y0, z0);
    const long long nlo = max((long long)x0.lo, yz.lo), nhi = min((long long)x0.hi, yz.hi);
    if (nlo > nhi) return E_FAIL;
    x.lo = (int)nlo; x.hi = (int)nhi;
    if (!stage<0>(c, us, t.xv, t.xo, x0, x.lo, x.hi)) return E_FAIL;
    const IV px = ro.x ? x0 : x;  // a multi-term Sum operand is checked, not narrowed (term/sum.rs:62-69)
    if ((long long)px.hi < yz.lo || yz.hi < (long long)px.lo) return E_FAIL;
    return (yz.lo == yz.hi && px.lo == px.hi) ? E_ENTAILED : E_UNKNOWN;
  }
  if (kind == T_EQ) {
    if (!prop_eq(x, y, z, ro)) return E_FAIL;
    s = sub_eq(x, y, z);
  } else if (kind == T_GREATER) {
    if (!prop_greater(x, y, z, 1)) return E_FAIL;
    undo_ro(ro, x, y, z, x0, y0, z0);
    s = sub_greater(x, y, z, 1);
  } else {
    if (!prop_less(x, y, z, 1)) return E_FAIL;
    undo_ro(ro, x, y, z, x0, y0, z0);
    s = sub_less(x, y, z, 1);
  }
  if (!stage_tri(c, us, t, x0, y0, z0, x, y, z)) return E_FAIL;
  return s < 0 ? E_FAIL : (s > 0 ? E_ENTAILED : E_UNKNOWN);
}
__device__ __forceinline__ bool ter_is_noop(unsigned kind, IV x, IV y, IV z) {
  if (kind == T_MUL) {
    const IV64 yz = mul_iv(y, z);
    return (long long)x.lo >= yz.lo && (long long)x.hi <= yz.hi && !(yz.lo == yz.hi && x.lo == x.hi);
  }
  if (kind == T_EQ)
    return x.lo >= y.lo + z.lo && y.hi <= x.hi - z.lo && z.hi <= x.hi - y.lo &&
           x.hi <= y.hi + z.hi && y.lo >= x.lo - z.hi && z.lo >= x.lo - y.hi && sub_eq(x, y, z) == 0;
  if (kind == T_GREATER)
    return x.lo >= y.lo + z.lo + 1 && y.hi <= x.hi - z.lo - 1 && z.hi <= x.hi - y.lo - 1 &&
           sub_greater(x, y, z, 1) == 0;
  return x.hi <= y.hi + z.hi - 1 && y.lo >= x.lo - z.hi + 1 && z.lo >= x.lo - y.hi + 1 && sub_less(x, y, z, 1) == 0;
}

// --- 2-way disjunction of XEqYPlusZ (logic/disjunction.rs:77-116) ---------------------------
template <bool SMEM>
__device__ __forceinline__ Eval eval_dj(const Ctx& c, int4 q0, int4 q1, int4 q2, UpdSet& us) {
  Tri a{q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
  Tri b{q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
  const IV ax0 = rd<SMEM>(c, a.xv, a.xo), ay0 = rd<SMEM>(c, a.yv, a.yo), az0 = rd<SMEM>(c, a.zv, a.zo);
  const IV bx0 = rd<SMEM>(c, b.xv, b.xo), by0 = rd<SMEM>(c, b.yv, b.yo), bz0 = rd<SMEM>(c, b.zv, b.zo);
  int sa = sub_eq(ax0, ay0, az0), sb = sub_eq(bx0, by0, bz0);
  if (sa > 0 || sb > 0) return E_ENTAILED;          // disjunction.rs:102: propagate -> true
  if (sa < 0 && sb < 0) return E_FAIL;              // disjunction.rs:110-111
  if (sa == 0 && sb == 0) return E_UNKNOWN;         // disjunction.rs:112-114
  if (sa < 0) {                                     // disjunction.rs:108-109
    IV x = bx0, y = by0, z = bz0;
    if (!prop_eq(x, y, z, tri_ro(b)) || !stage_tri(c, us, b, bx0, by0, bz0, x, y, z)) return E_FAIL;
    sb = sub_eq(x, y, z);
    return sb < 0 ? E_FAIL : (sb > 0 ? E_ENTAILED : E_UNKNOWN);
  }
  IV x = ax0, y = ay0, z = az0;
  if (!prop_eq(x, y, z, tri_ro(a)) || !stage_tri(c, us, a, ax0, ay0, az0, x, y, z)) return E_FAIL;
  sa = sub_eq(x, y, z);
  return sa < 0 ? E_FAIL : (sa > 0 ? E_ENTAILED : E_UNKNOWN);
}
template <bool SMEM>
__device__ __forceinline__ bool dj_is_noop(const Ctx& c, int4 q0, int4 q1, int4 q2) {
  IV ax = rd<SMEM>(c, q0.x, q0.y), ay = rd<SMEM>(c, q0.z, q0.w), az = rd<SMEM>(c, q1.x, q1.y);
  IV bx = rd<SMEM>(c, q1.z, q1.w), by = rd<SMEM>(c, q2.x, q2.y), bz = rd<SMEM>(c, q2.z, q2.w);
  return sub_eq(ax, ay, az) == 0 && sub_eq(bx, by, bz) == 0;
}

// The out-of-line slow paths: full propagate + is_subsumed of one propagator, with the
// store bookkeeping of store.rs:166-207 (failure flag, unlink of entailed propagators).
// A failed propagate applies nothing (the node is discarded anyway).
__device__ __forceinline__ void finish_eval(const Ctx& c, Eval r, const UpdSet& us, unsigned fam, int slot) {
  if (r == E_FAIL) { set_failed(c); return; }
  apply_updates(c, us);
  if (r == E_ENTAILED && c.bookkeep) deactivate(c, c.P->fam[fam].active, fam, slot);
}
__device__ __noinline__ void eval_full_bin(const Ctx& c, int slot, int4 d, IV x, IV y) {
  UpdSet us;
  upd_clear(us);
  finish_eval(c, eval_bin(c, d, x, y, us), us, F_BIN, slot);
}
__device__ __noinline__ void eval_full_ter(const Ctx& c, int slot, int4 a, int2 b, IV x, IV y, IV z) {
  UpdSet us;
  upd_clear(us);
  finish_eval(c, eval_ter(c, a, b, x, y, z, us), us, F_TER, slot);
}
template <bool SMEM>
__device__ __noinline__ void eval_full_dj(const Ctx& c, int slot, int4 q0, int4 q1, int4 q2) {
  UpdSet us;
  upd_clear(us);
  finish_eval(c, eval_dj<SMEM>(c, q0, q1, q2, us), us, F_DJ, slot);
}
// Any family, operands read here (posted propagators, tail).
template <bool SMEM>
__device__ __forceinline__ void eval_full(const Ctx& c, unsigned fam, int slot, int4 q0, int4 q1, int4 q2) {
  if (fam == F_BIN) {
    eval_full_bin(c, slot, q0, rd<SMEM>(c, dec_var28((unsigned)q0.x), q0.y), rd<SMEM>(c, q0.z, q0.w));
  }
#ifndef PCP_BIN_ONLY
  else if (fam == F_TER) {
    eval_full_ter(c, slot, q0, make_int2(q1.x, q1.y), rd<SMEM>(c, dec_var28((unsigned)q0.x), q0.y), rd<SMEM>(c, q0.z, q0.w),
                  rd<SMEM>(c, q1.x, q1.y));
  } else {
    eval_full_dj<SMEM>(c, slot, q0, q1, q2);
  }
#endif
}

// Gather one descriptor from global memory (worklist expansion, tail).
__device__ __forceinline__ void load_desc(const Family& f, unsigned fam, int slot, int4& q0, int4& q1, int4& q2) {
  q1 = make_int4(0, 0, 0, 0);
  q2 = q1;
  const int4* desc = slot < f.n_static ? f.desc : f.tdesc;  // (a shared static part keeps the tail apart)
  if (fam == F_BIN) {
    q0 = __ldg(&desc[slot]);
  } else if (fam == F_TER) {
    q0 = __ldg(&desc[slot]);
    int2 b = __ldg(&(slot < f.n_static ? f.descB : f.tdescB)[slot]);
    q1.x = b.x; q1.y = b.y;
  } else {
    const int4* q = &desc[3 * (size_t)slot];
    q0 = __ldg(q); q1 = __ldg(q + 1); q2 = __ldg(q + 2);
  }
}
// Evaluate a gathered propagator: cheap no-op test inline, everything else out of line.
template <bool SMEM>
__device__ __forceinline__ void eval_loaded(const Ctx& c, unsigned fam, int slot, int4 q0, int4 q1, int4 q2) {
  if (fam == F_BIN) {
    IV x = rd<SMEM>(c, dec_var28((unsigned)q0.x), q0.y), y = rd<SMEM>(c, q0.z, q0.w);
#ifdef PCP_SET
    if (!bin_is_noop_set(c, q0, x, y)) eval_full_bin(c, slot, q0, x, y);
#else
    if (!bin_is_noop((unsigned)q0.x >> 28, x, y)) eval_full_bin(c, slot, q0, x, y);
#endif
  }
#ifndef PCP_BIN_ONLY
  else if (fam == F_TER) {
    IV x = rd<SMEM>(c, dec_var28((unsigned)q0.x), q0.y), y = rd<SMEM>(c, q0.z, q0.w), z = rd<SMEM>(c, q1.x, q1.y);
    if (!ter_is_noop((unsigned)q0.x >> 28, x, y, z)) eval_full_ter(c, slot, q0, make_int2(q1.x, q1.y), x, y, z);
  } else {
    if (!dj_is_noop<SMEM>(c, q0, q1, q2)) eval_full_dj<SMEM>(c, slot, q0, q1, q2);
  }
#endif
}
template <bool SMEM>
__device__ __forceinline__ void eval_ref(const Ctx& c, unsigned fam, int slot) {
  int4 q0, q1, q2;
  load_desc(c.P->fam[fam], fam, slot, q0, q1, q2);
  eval_loaded<SMEM>(c, fam, slot, q0, q1, q2);
}

__device__ __forceinline__ bool is_active(const Family& f, int slot) {
  return (__ldcg(&f.active[slot >> 5]) >> (slot & 31)) & 1u;
}

