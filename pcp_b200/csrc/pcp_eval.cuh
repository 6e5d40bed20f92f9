// pcp_eval.cuh -- per-propagator evaluation (propagate + is_subsumed) on the device.
// Included by pcp_device.cuh after Params / the memory helpers.
//
// Every propagator is evaluated in two steps: a *pure* step that reads the operand views
// and computes the narrowed bounds exactly as the reference's `propagate` would (file:line
// at each function), staging at most one update per operand; and `apply_updates`, which
// issues all staged atomicMax/atomicMin together (their L2 round trips overlap), then
// queues the variables that really changed on the dirty worklist with one warp-aggregated
// reservation.  A cheap inline "would this evaluation change anything?" test keeps the
// common case of a sweep free of calls.
#pragma once

namespace pcpd {

struct IV { int lo, hi; };

// Per-thread context of one fixpoint launch (read-mostly; counters live in registers).
struct Ctx {
  const Params* P;
  int2* sdom;             // shared-memory snapshot (or nullptr)
  unsigned next_epoch;    // stamp for "dirty in the next iteration"
  int next_buf;           // dirty list written in this iteration
  bool local;             // updates go to the shared-memory snapshot only (posted props, per CTA)
  bool mark_dirty;        // narrowed variables enter the worklist
  bool bookkeep;          // entailed propagators are deactivated + trailed
  int* flags;             // shared: [0] this CTA queued a dirty variable, [1] saw a failure
  // solo mode (a short cascade run by CTA 0 alone, see solo_iterations): the snapshot is the
  // authoritative copy, updated with shared-memory atomics and written through to HBM;
  // the next worklist is a bit set + short list in shared memory
  bool mirror;            // row-local fixpoint: every update is also applied to this CTA's snapshot
  bool solo;
  unsigned* solo_next_bits;
  int* solo_next_list;
  int* solo_next_cnt;
  int solo_cap;
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
      int2 d = SMEM ? c.sdom[term.x] : ldcg_dom(&P.dom[term.x]);
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
  int2 d = SMEM ? c.sdom[var] : ldcg_dom(&c.P->dom[var]);
  return IV{d.x + off, d.y + off};   // Addition (term/addition.rs:93-101)
}

// ---------------------------------------------------------------------------------------
// staged updates: variable/store.rs:151-166 through term/addition.rs:80-90 /
// term/constant.rs:43-53.  All bounds are in view space.
// ---------------------------------------------------------------------------------------
struct Upd { int var, off, cur_lo, cur_hi, nlo, nhi; };
struct UpdSet {
  Upd u[3];
  int n;
};

// Stage "view <- [nlo, nhi]" (already intersected with `cur`).  Returns false when the new
// domain is empty (the reference's update -> false); a Constant is never written but fails
// when it falls outside (nlo > nhi covers it because cur is the singleton).
__device__ __forceinline__ bool stage(UpdSet& us, int var, int off, IV cur, int nlo, int nhi) {
  if (nlo > nhi) return false;
  if (var >= 0 && (nlo > cur.lo || nhi < cur.hi)) {
    Upd& r = us.u[us.n++];
    r.var = var; r.off = off; r.cur_lo = cur.lo; r.cur_hi = cur.hi; r.nlo = nlo; r.nhi = nhi;
  }
  return true;
}

__device__ __forceinline__ void apply_updates(const Ctx& c, const UpdSet& us) {
  if (us.n == 0) return;
  if (c.local) {
    for (int i = 0; i < us.n; ++i) c.sdom[us.u[i].var] = make_int2(us.u[i].nlo - us.u[i].off, us.u[i].nhi - us.u[i].off);
    return;
  }
  const Params& P = *c.P;
  if (c.solo) {
    for (int i = 0; i < us.n; ++i) {
      const Upd& r = us.u[i];
      const int nl = r.nlo - r.off, nh = r.nhi - r.off;
      int2* s = &c.sdom[r.var];
      int2* d = &P.dom[r.var];
      bool ch = false;
      int lo_now = r.cur_lo - r.off, hi_now = r.cur_hi - r.off;
      if (r.nlo > r.cur_lo) {
        int old = atomicMax(&s->x, nl);     // shared-memory atomic: the authoritative copy
        if (old < nl) { ch = true; atomicMax(&d->x, nl); }  // write-through, result unused (RED)
        lo_now = max(old, nl);
      }
      if (r.nhi < r.cur_hi) {
        int old = atomicMin(&s->y, nh);
        if (old > nh) { ch = true; atomicMin(&d->y, nh); }
        hi_now = min(old, nh);
      }
      if (!ch) continue;
      // the other bound may have moved under us: re-read the authoritative copy
      const int now_lo = *(volatile int*)&s->x, now_hi = *(volatile int*)&s->y;
      if (lo_now > hi_now || now_lo > now_hi) set_failed(c);
      const unsigned bit = 1u << (r.var & 31);
      if (!(atomicOr(&c.solo_next_bits[r.var >> 5], bit) & bit)) {
        int idx = atomicAdd(c.solo_next_cnt, 1);
        if (idx < c.solo_cap) c.solo_next_list[idx] = r.var;
      }
    }
    return;
  }
  // 1. all atomics in flight together (plus a plain look at the dirty stamp: a variable that is
  //    already queued for the next iteration needs no exchange)
  int old_lo[3], old_hi[3];
  unsigned seen[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < us.n) {
      const Upd& r = us.u[i];
      int2* d = &P.dom[r.var];
      old_lo[i] = r.nlo > r.cur_lo ? atomicMax(&d->x, r.nlo - r.off) : r.cur_lo - r.off;
      old_hi[i] = r.nhi < r.cur_hi ? atomicMin(&d->y, r.nhi - r.off) : r.cur_hi - r.off;
      seen[i] = c.mark_dirty ? __ldcg(&P.dirty_stamp[r.var]) : 0u;
      if (c.mirror) {  // keep the CTA's snapshot current for the next local round
        if (r.nlo > r.cur_lo) atomicMax(&c.sdom[r.var].x, r.nlo - r.off);
        if (r.nhi < r.cur_hi) atomicMin(&c.sdom[r.var].y, r.nhi - r.off);
      }
    }
  }
  // 2. which variables did this thread really narrow?
  bool ch[3] = {false, false, false};
  bool fail = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < us.n) {
      const Upd& r = us.u[i];
      int nl = r.nlo - r.off, nh = r.nhi - r.off;
      ch[i] = (r.nlo > r.cur_lo && old_lo[i] < nl) || (r.nhi < r.cur_hi && old_hi[i] > nh);
      // a domain emptied by concurrent updates this thread cannot see is caught when the
      // variable is refreshed from the worklist in the next iteration
      fail |= max(old_lo[i], nl) > min(old_hi[i], nh);
    }
  }
  if (fail) set_failed(c);
  if (!c.mark_dirty) return;
  // 3. queue them: stamps first (dedup across the grid), then one reservation per warp
  bool nw[3] = {false, false, false};
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < us.n && ch[i] && seen[i] != c.next_epoch)
      nw[i] = atomicExch(&P.dirty_stamp[us.u[i].var], c.next_epoch) != c.next_epoch;
  if (ch[0] | ch[1] | ch[2]) c.flags[0] = 1;
  const unsigned m = __activemask();
  const unsigned b0 = __ballot_sync(m, nw[0]), b1 = __ballot_sync(m, nw[1]), b2 = __ballot_sync(m, nw[2]);
  const int total = __popc(b0) + __popc(b1) + __popc(b2);
  if (total == 0) return;
  const int leader = __ffs(m) - 1;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == leader) base = atomicAdd(&P.ctl->dirty_cnt[c.next_buf], total);
  base = __shfl_sync(m, base, leader);
  int* list = P.dirty_list + (size_t)c.next_buf * P.V;
  const unsigned lt = lanemask_lt();
  if (nw[0]) list[base + __popc(b0 & lt)] = us.u[0].var;
  if (nw[1]) list[base + __popc(b0) + __popc(b1 & lt)] = us.u[1].var;
  if (nw[2]) list[base + __popc(b0) + __popc(b1) + __popc(b2 & lt)] = us.u[2].var;
}

// Single update (n-ary propagators write their operands one by one).
__device__ __forceinline__ bool tighten(const Ctx& c, int var, int off, IV cur, int nlo, int nhi) {
  UpdSet us;
  us.n = 0;
  if (!stage(us, var, off, cur, nlo, nhi)) return false;
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
template <bool SMEM>
__device__ __forceinline__ Eval eval_bin(const Ctx& c, int4 d, UpdSet& us) {
  unsigned kind = (unsigned)d.x >> 28;
  int xv = dec_var28((unsigned)d.x), xo = d.y, yv = d.z, yo = d.w;
  const IV x = rd<SMEM>(c, xv, xo), y = rd<SMEM>(c, yv, yo);
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
    if (!stage(us, yv, yo, y, ny.lo, ny.hi)) return E_FAIL;
    if (!stage(us, xv, xo, x, nx.lo, nx.hi)) return E_FAIL;
  } else if (kind == B_LESS) {  // cmp/x_less_y.rs:101-108
    nx.hi = min(x.hi, y.hi - 1);
    if (!stage(us, xv, xo, x, nx.lo, nx.hi)) return E_FAIL;
    ny.lo = max(y.lo, x.lo + 1);
    if (!stage(us, yv, yo, y, ny.lo, ny.hi)) return E_FAIL;
  } else {  // B_EQ: cmp/x_eq_y.rs:102-107
    nx.lo = ny.lo = max(x.lo, y.lo);
    nx.hi = ny.hi = min(x.hi, y.hi);
    if (!stage(us, xv, xo, x, nx.lo, nx.hi)) return E_FAIL;
    if (!stage(us, yv, yo, y, ny.lo, ny.hi)) return E_FAIL;
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
__device__ __forceinline__ bool stage_tri(UpdSet& us, const Tri& t, IV x0, IV y0, IV z0, IV x, IV y, IV z) {
  return stage(us, t.xv, t.xo, x0, x.lo, x.hi) && stage(us, t.yv, t.yo, y0, y.lo, y.hi) &&
         stage(us, t.zv, t.zo, z0, z.lo, z.hi);
}

template <bool SMEM>
__device__ __forceinline__ Eval eval_ter(const Ctx& c, int4 a, int2 b, UpdSet& us) {
  unsigned kind = (unsigned)a.x >> 28;
  Tri t{dec_var28((unsigned)a.x), a.y, a.z, a.w, b.x, b.y};
  const IV x0 = rd<SMEM>(c, t.xv, t.xo), y0 = rd<SMEM>(c, t.yv, t.yo), z0 = rd<SMEM>(c, t.zv, t.zo);
  IV x = x0, y = y0, z = z0;
  const TriRo ro = tri_ro(t);
  int s;
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
  if (!stage_tri(us, t, x0, y0, z0, x, y, z)) return E_FAIL;
  return s < 0 ? E_FAIL : (s > 0 ? E_ENTAILED : E_UNKNOWN);
}
__device__ __forceinline__ bool ter_is_noop(unsigned kind, IV x, IV y, IV z) {
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
    if (!prop_eq(x, y, z, tri_ro(b)) || !stage_tri(us, b, bx0, by0, bz0, x, y, z)) return E_FAIL;
    sb = sub_eq(x, y, z);
    return sb < 0 ? E_FAIL : (sb > 0 ? E_ENTAILED : E_UNKNOWN);
  }
  IV x = ax0, y = ay0, z = az0;
  if (!prop_eq(x, y, z, tri_ro(a)) || !stage_tri(us, a, ax0, ay0, az0, x, y, z)) return E_FAIL;
  sa = sub_eq(x, y, z);
  return sa < 0 ? E_FAIL : (sa > 0 ? E_ENTAILED : E_UNKNOWN);
}
template <bool SMEM>
__device__ __forceinline__ bool dj_is_noop(const Ctx& c, int4 q0, int4 q1, int4 q2) {
  IV ax = rd<SMEM>(c, q0.x, q0.y), ay = rd<SMEM>(c, q0.z, q0.w), az = rd<SMEM>(c, q1.x, q1.y);
  IV bx = rd<SMEM>(c, q1.z, q1.w), by = rd<SMEM>(c, q2.x, q2.y), bz = rd<SMEM>(c, q2.z, q2.w);
  return sub_eq(ax, ay, az) == 0 && sub_eq(bx, by, bz) == 0;
}

// The out-of-line slow path: full propagate + is_subsumed of one propagator, with the
// store bookkeeping of store.rs:166-207 (failure flag, unlink of entailed propagators).
// A failed propagate applies nothing (the node is discarded anyway).
template <bool SMEM>
__device__ __noinline__ void eval_full(const Ctx& c, unsigned fam, int slot, int4 q0, int4 q1, int4 q2) {
  UpdSet us;
  us.n = 0;
  Eval r;
  if (fam == F_BIN) r = eval_bin<SMEM>(c, q0, us);
  else if (fam == F_TER) r = eval_ter<SMEM>(c, q0, make_int2(q1.x, q1.y), us);
  else r = eval_dj<SMEM>(c, q0, q1, q2, us);
  if (r == E_FAIL) { set_failed(c); return; }
  apply_updates(c, us);
  if (r == E_ENTAILED && c.bookkeep) deactivate(c, c.P->fam[fam].active, fam, slot);
}

// Gather one descriptor from global memory (worklist expansion, tail).
__device__ __forceinline__ void load_desc(const Family& f, unsigned fam, int slot, int4& q0, int4& q1, int4& q2) {
  q1 = make_int4(0, 0, 0, 0);
  q2 = q1;
  if (fam == F_BIN) {
    q0 = __ldg(&f.desc[slot]);
  } else if (fam == F_TER) {
    q0 = __ldg(&f.desc[slot]);
    int2 b = __ldg(&f.descB[slot]);
    q1.x = b.x; q1.y = b.y;
  } else {
    const int4* q = &f.desc[3 * (size_t)slot];
    q0 = __ldg(q); q1 = __ldg(q + 1); q2 = __ldg(q + 2);
  }
}
// Evaluate a gathered propagator: cheap no-op test inline, everything else out of line.
template <bool SMEM>
__device__ __forceinline__ void eval_loaded(const Ctx& c, unsigned fam, int slot, int4 q0, int4 q1, int4 q2) {
  if (fam == F_BIN) {
    IV x = rd<SMEM>(c, dec_var28((unsigned)q0.x), q0.y), y = rd<SMEM>(c, q0.z, q0.w);
    if (bin_is_noop((unsigned)q0.x >> 28, x, y)) return;
  } else if (fam == F_TER) {
    IV x = rd<SMEM>(c, dec_var28((unsigned)q0.x), q0.y), y = rd<SMEM>(c, q0.z, q0.w), z = rd<SMEM>(c, q1.x, q1.y);
    if (ter_is_noop((unsigned)q0.x >> 28, x, y, z)) return;
  } else {
    if (dj_is_noop<SMEM>(c, q0, q1, q2)) return;
  }
  eval_full<SMEM>(c, fam, slot, q0, q1, q2);
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

}  // namespace pcpd
