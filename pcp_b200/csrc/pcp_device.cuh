// pcp_device.cuh -- sm_100a device side of the propagation engine.
//
// One persistent kernel (`pcp_fixpoint_kernel`) runs the whole propagation fixpoint of a
// search node -- the device equivalent of Store::consistency
// (reference src/libpcp/propagation/store.rs:247-257): one CTA per SM, a device-wide
// barrier between iterations, quiescence detected by the last CTA to arrive.
//
//   iteration 0   every active propagator is evaluated once (store.rs:144-149 schedules
//                 all active propagators): a streaming scan over the per-family descriptor
//                 arrays with 128-bit coalesced loads, domains gathered from a
//                 shared-memory snapshot (small V) or from L2 (large V).
//   iteration k   only propagators adjacent to variables that changed in iteration k-1
//                 are re-evaluated (store.rs:191-198 `react`): the dirty variables sit in
//                 a warp-aggregated worklist, their rows of the static var->propagator
//                 CSR (the reactor, reactors/indexed_deps.rs:23-27) are expanded warp by
//                 warp, and a per-propagator epoch stamp plays the role of RelaxedFifo's
//                 `inside_queue` bit set (schedulers/relaxed_fifo.rs:42-48).
//
// Updates are monotone atomics (atomicMax on lo, atomicMin on hi), so the chaotic
// iteration converges to the same greatest fixpoint as the reference FIFO (SURVEY 8a,
// "Parity theorem").  Integer bound arithmetic only: no tensor cores, no floating point.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcpd {

constexpr int kThreads = 1024;          // one CTA per SM
constexpr int kWarps = kThreads / 32;
constexpr int kUnroll = 4;              // descriptor loads in flight per thread
constexpr unsigned kConstVar28 = 0x0FFFFFFFu;

// family tags inside adjacency / trail references (top 3 bits)
enum Fam : unsigned { F_BIN = 0, F_TER = 1, F_DJ = 2, F_NARY = 3 };
constexpr unsigned kSlotMask = 0x1FFFFFFFu;
__host__ __device__ inline unsigned make_ref(unsigned fam, unsigned slot) { return (fam << 29) | slot; }

// kinds inside a family (top 4 bits of descriptor word 0)
enum BinKind : unsigned { B_LESS = 0, B_NEQ = 1, B_EQ = 2 };
enum TerKind : unsigned { T_GREATER = 0, T_LESS = 1, T_EQ = 2 };

enum Decision : unsigned { D_CONTINUE = 0, D_FIXPOINT = 1, D_FAILED = 2, D_ITER_CAP = 3 };

struct Control {
  unsigned bar_count;
  unsigned bar_gen;
  unsigned decision[2];
  unsigned epoch;        // next unused epoch (stamps < epoch are stale)
  int failed;
  unsigned trail_cnt;    // number of deactivated (entailed) propagators on the trail
  int dirty_cnt[3];
  unsigned long long propagations;  // cumulative
  unsigned iterations;   // of the last launch
  unsigned last_decision;
};

// Header copied back to the host together with the domains (one D2H copy per node).
struct Result {
  int failed;
  unsigned trail_cnt;
  unsigned iterations;
  unsigned epoch;
  unsigned long long propagations;
  unsigned decision;
  unsigned pad[9];
};
static_assert(sizeof(Result) == 64, "Result header is 64 bytes");

struct Family {
  const int4* desc;     // BIN: 1 int4/prop; TER: int4 (x,y) plane; DJ: 3 int4/prop
  const int2* descB;    // TER: (z) plane
  uint32_t* active;     // bit set, 1 = active (propagation/store.rs:34)
  uint32_t* stamp;      // epoch of the last worklist evaluation
  int n;                // allocated propagators
  int n_static;         // [0, n_static) are covered by the CSR; [n_static, n) is the tail
};

struct Params {
  Result* result;
  int2* dom;            // interval per variable: (lo, hi)
  int V;
  int smem_dom;         // 1: domains are staged in shared memory
  Family bin, ter, dj;
  const int* nary_ptr;  // CSR of n-ary Distinct operands
  const int2* nary_ops;
  uint32_t* nary_active;
  int n_nary;
  int nary_max_k;
  const int* adj_ptr;   // reactor: var -> propagator refs
  const uint32_t* adj;
  int* dirty_list;      // 3 x V
  uint32_t* dirty_stamp;
  uint32_t* trail;
  Control* ctl;
  int full_sweep;       // 1: schedule every active propagator first (store.rs:144-149)
  unsigned max_iterations;
};

// ---------------------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int2 ldcg_dom(const int2* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

struct IV { int lo, hi; };

// Per-thread context of one fixpoint launch.
struct Ctx {
  const Params* P;
  int2* sdom;             // shared-memory snapshot (or nullptr)
  unsigned next_epoch;    // stamp for "dirty in the next iteration"
  int next_buf;           // dirty list written in this iteration
  unsigned nprop;         // propagations executed by this thread
  bool count;             // false for redundant (all-CTA) tail evaluations
};

// warp-aggregated append to the dirty-variable worklist
__device__ __forceinline__ void push_dirty(Ctx& c, int v) {
  const Params& P = *c.P;
  if (atomicExch(&P.dirty_stamp[v], c.next_epoch) == c.next_epoch) return;  // already queued
  unsigned m = __activemask();
  int leader = __ffs(m) - 1;
  int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == leader) base = atomicAdd(&P.ctl->dirty_cnt[c.next_buf], __popc(m));
  base = __shfl_sync(m, base, leader);
  P.dirty_list[c.next_buf * P.V + base + __popc(m & lanemask_lt())] = v;
}

// warp-aggregated append to the entailment trail + clear of the active bit
__device__ __forceinline__ void deactivate(Ctx& c, const Family& f, unsigned fam, int slot) {
  unsigned bit = 1u << (slot & 31);
  unsigned old = atomicAnd(&f.active[slot >> 5], ~bit);
  if (!(old & bit)) return;
  unsigned m = __activemask();
  int leader = __ffs(m) - 1;
  int lane = threadIdx.x & 31;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(&c.P->ctl->trail_cnt, (unsigned)__popc(m));
  base = __shfl_sync(m, base, leader);
  c.P->trail[base + __popc(m & lanemask_lt())] = make_ref(fam, (unsigned)slot);
}

__device__ __forceinline__ void set_failed(Ctx& c) { c.P->ctl->failed = 1; }

// Domain readers.  `SMEM` reads the CTA's snapshot, otherwise L2 (ld.global.cg).
template <bool SMEM>
__device__ __forceinline__ IV rd(const Ctx& c, int var, int off) {
  if (var < 0) return IV{off, off};  // Constant (term/constant.rs:55-63)
  int2 d = SMEM ? c.sdom[var] : ldcg_dom(&c.P->dom[var]);
  return IV{d.x + off, d.y + off};   // Addition (term/addition.rs:93-101)
}

// Monotone update of one view to [nlo, nhi] (already intersected with `cur`):
// variable/store.rs:151-166 through term/addition.rs:80-90 / term/constant.rs:43-53.
// Returns false when the new domain is empty (update -> false).
__device__ __forceinline__ bool tighten(Ctx& c, int var, int off, IV cur, int nlo, int nhi) {
  if (nlo > nhi) return false;
  if (var >= 0) {
    int2* d = &c.P->dom[var];
    bool ch = false;
    if (nlo > cur.lo) ch |= atomicMax(&d->x, nlo - off) < nlo - off;
    if (nhi < cur.hi) ch |= atomicMin(&d->y, nhi - off) > nhi - off;
    if (ch) {
      int2 now = ldcg_dom(d);
      if (now.x > now.y) set_failed(c);
      push_dirty(c, var);
    }
  }
  return true;
}

// Result of evaluating one propagator: propagate + is_subsumed (store.rs:177-183).
enum Eval : int { E_FAIL = -1, E_UNKNOWN = 0, E_ENTAILED = 1 };

// --- binary family: XLessY / XNeqY / XEqY ---------------------------------------------------
template <bool SMEM>
__device__ __forceinline__ Eval eval_bin(Ctx& c, int4 d) {
  unsigned w0 = (unsigned)d.x;
  unsigned kind = w0 >> 28;
  int xv = (w0 & kConstVar28) == kConstVar28 ? -1 : (int)(w0 & kConstVar28);
  int xo = d.y, yv = d.z, yo = d.w;
  IV x = rd<SMEM>(c, xv, xo), y = rd<SMEM>(c, yv, yo);
  if (kind == B_NEQ) {  // cmp/x_neq_y.rs:82-93 + Interval::difference
    IV nx = x, ny = y;
    if (x.lo == x.hi) {
      if (ny.lo == x.lo) ny.lo++; else if (ny.hi == x.lo) ny.hi--;
    } else if (y.lo == y.hi) {
      if (nx.lo == y.lo) nx.lo++; else if (nx.hi == y.lo) nx.hi--;
    }
    if (!tighten(c, yv, yo, y, ny.lo, ny.hi)) return E_FAIL;
    if (!tighten(c, xv, xo, x, nx.lo, nx.hi)) return E_FAIL;
    // !XEqY::is_subsumed (x_neq_y.rs:71-73, x_eq_y.rs:84-93)
    return (nx.hi < ny.lo || ny.hi < nx.lo) ? E_ENTAILED : E_UNKNOWN;
  } else if (kind == B_LESS) {  // cmp/x_less_y.rs:101-108
    int nxhi = min(x.hi, y.hi - 1);
    if (!tighten(c, xv, xo, x, x.lo, nxhi)) return E_FAIL;
    int nylo = max(y.lo, x.lo + 1);
    if (!tighten(c, yv, yo, y, nylo, y.hi)) return E_FAIL;
    return nxhi < nylo ? E_ENTAILED : E_UNKNOWN;  // x_less_y.rs:84-91
  } else {  // B_EQ: cmp/x_eq_y.rs:102-107
    int lo = max(x.lo, y.lo), hi = min(x.hi, y.hi);
    if (!tighten(c, xv, xo, x, lo, hi)) return E_FAIL;
    if (!tighten(c, yv, yo, y, lo, hi)) return E_FAIL;
    return lo == hi ? E_ENTAILED : E_UNKNOWN;     // x_eq_y.rs:84-93
  }
}

// --- ternary family ------------------------------------------------------------------------
struct Tri { int xv, xo, yv, yo, zv, zo; };

// XGreaterYPlusZ::propagate on local copies (x_greater_y_plus_z.rs:106-119); `strict`=1 for
// x > y+z, 0 for x >= y+z (the Addition(x,1) of cmp/mod.rs:73).
template <bool SMEM>
__device__ __forceinline__ bool prop_greater(Ctx& c, const Tri& t, IV& x, IV& y, IV& z, int strict) {
  int nxlo = max(x.lo, y.lo + z.lo + strict);
  int nyhi = min(y.hi, x.hi - z.lo - strict);
  int nzhi = min(z.hi, x.hi - y.lo - strict);
  if (!tighten(c, t.xv, t.xo, x, nxlo, x.hi)) return false;
  if (!tighten(c, t.yv, t.yo, y, y.lo, nyhi)) return false;
  if (!tighten(c, t.zv, t.zo, z, z.lo, nzhi)) return false;
  x.lo = nxlo; y.hi = nyhi; z.hi = nzhi;
  return true;
}
// XLessYPlusZ::propagate (x_less_y_plus_z.rs:106-120)
template <bool SMEM>
__device__ __forceinline__ bool prop_less(Ctx& c, const Tri& t, IV& x, IV& y, IV& z, int strict) {
  int nxhi = min(x.hi, y.hi + z.hi - strict);
  int nylo = max(y.lo, x.lo - z.hi + strict);
  int nzlo = max(z.lo, x.lo - y.hi + strict);
  if (!tighten(c, t.xv, t.xo, x, x.lo, nxhi)) return false;
  if (!tighten(c, t.yv, t.yo, y, nylo, y.hi)) return false;
  if (!tighten(c, t.zv, t.zo, z, nzlo, z.hi)) return false;
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
// XEqYPlusZ::propagate = geq then leq re-reading the store (x_eq_y_plus_z.rs:79-81)
template <bool SMEM>
__device__ __forceinline__ bool prop_eq(Ctx& c, const Tri& t, IV& x, IV& y, IV& z) {
  return prop_greater<SMEM>(c, t, x, y, z, 0) && prop_less<SMEM>(c, t, x, y, z, 0);
}

template <bool SMEM>
__device__ __forceinline__ Eval eval_ter(Ctx& c, int4 a, int2 b) {
  unsigned w0 = (unsigned)a.x;
  unsigned kind = w0 >> 28;
  Tri t;
  t.xv = (w0 & kConstVar28) == kConstVar28 ? -1 : (int)(w0 & kConstVar28);
  t.xo = a.y; t.yv = a.z; t.yo = a.w; t.zv = b.x; t.zo = b.y;
  IV x = rd<SMEM>(c, t.xv, t.xo), y = rd<SMEM>(c, t.yv, t.yo), z = rd<SMEM>(c, t.zv, t.zo);
  int s;
  if (kind == T_EQ) {
    if (!prop_eq<SMEM>(c, t, x, y, z)) return E_FAIL;
    s = sub_eq(x, y, z);
  } else if (kind == T_GREATER) {
    if (!prop_greater<SMEM>(c, t, x, y, z, 1)) return E_FAIL;
    s = sub_greater(x, y, z, 1);
  } else {
    if (!prop_less<SMEM>(c, t, x, y, z, 1)) return E_FAIL;
    s = sub_less(x, y, z, 1);
  }
  return s < 0 ? E_FAIL : (s > 0 ? E_ENTAILED : E_UNKNOWN);
}

// --- 2-way disjunction of XEqYPlusZ (logic/disjunction.rs:77-116) ---------------------------
template <bool SMEM>
__device__ __forceinline__ Eval eval_dj(Ctx& c, int4 q0, int4 q1, int4 q2) {
  Tri a{q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
  Tri b{q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
  IV ax = rd<SMEM>(c, a.xv, a.xo), ay = rd<SMEM>(c, a.yv, a.yo), az = rd<SMEM>(c, a.zv, a.zo);
  IV bx = rd<SMEM>(c, b.xv, b.xo), by = rd<SMEM>(c, b.yv, b.yo), bz = rd<SMEM>(c, b.zv, b.zo);
  int sa = sub_eq(ax, ay, az), sb = sub_eq(bx, by, bz);
  if (sa > 0 || sb > 0) return E_ENTAILED;          // disjunction.rs:102: propagate -> true
  if (sa < 0 && sb < 0) return E_FAIL;              // disjunction.rs:110-111
  if (sa == 0 && sb == 0) return E_UNKNOWN;         // disjunction.rs:112-114
  if (sa < 0) {                                     // disjunction.rs:108-109
    if (!prop_eq<SMEM>(c, b, bx, by, bz)) return E_FAIL;
    sb = sub_eq(bx, by, bz);
    return sb < 0 ? E_FAIL : (sb > 0 ? E_ENTAILED : E_UNKNOWN);
  }
  if (!prop_eq<SMEM>(c, a, ax, ay, az)) return E_FAIL;
  sa = sub_eq(ax, ay, az);
  return sa < 0 ? E_FAIL : (sa > 0 ? E_ENTAILED : E_UNKNOWN);
}

__device__ __forceinline__ void finish(Ctx& c, Eval r, const Family& f, unsigned fam, int slot) {
  if (c.count) c.nprop++;
  if (r == E_FAIL) set_failed(c);
  else if (r == E_ENTAILED && c.count) deactivate(c, f, fam, slot);  // store.rs:200-207 unlink_prop
}

template <bool SMEM>
__device__ __forceinline__ void eval_ref(Ctx& c, unsigned fam, int slot) {
  const Params& P = *c.P;
  if (fam == F_BIN) {
    finish(c, eval_bin<SMEM>(c, __ldg(&P.bin.desc[slot])), P.bin, F_BIN, slot);
  } else if (fam == F_TER) {
    finish(c, eval_ter<SMEM>(c, __ldg(&P.ter.desc[slot]), __ldg(&P.ter.descB[slot])), P.ter, F_TER, slot);
  } else {
    const int4* q = &P.dj.desc[3 * (size_t)slot];
    finish(c, eval_dj<SMEM>(c, __ldg(q), __ldg(q + 1), __ldg(q + 2)), P.dj, F_DJ, slot);
  }
}

__device__ __forceinline__ bool is_active(const Family& f, int slot) {
  return (__ldcg(&f.active[slot >> 5]) >> (slot & 31)) & 1u;
}

// ---------------------------------------------------------------------------------------
// iteration 0: streaming sweep over one family's static range, kUnroll x 32 propagators per
// warp step (coalesced 128-bit descriptor loads, all issued before the first use).
// ---------------------------------------------------------------------------------------
template <bool SMEM, unsigned FAM>
__device__ __forceinline__ void sweep_family(Ctx& c, const Family& f) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarps;
  const int n = f.n_static;
  constexpr int kChunk = 32 * kUnroll;
  for (long long base = warp * kChunk; base < n; base += nwarps * kChunk) {
    uint32_t aw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      long long p = base + u * 32;
      aw[u] = p < n ? __ldcg(&f.active[p >> 5]) : 0u;
    }
    if (FAM == F_BIN) {
      int4 d[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        int p = (int)base + u * 32 + lane;
        if (p < n && ((aw[u] >> lane) & 1u)) d[u] = __ldg(&f.desc[p]);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        int p = (int)base + u * 32 + lane;
        if (p < n && ((aw[u] >> lane) & 1u)) finish(c, eval_bin<SMEM>(c, d[u]), f, F_BIN, p);
      }
    } else if (FAM == F_TER) {
      int4 a[kUnroll];
      int2 b[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        int p = (int)base + u * 32 + lane;
        if (p < n && ((aw[u] >> lane) & 1u)) { a[u] = __ldg(&f.desc[p]); b[u] = __ldg(&f.descB[p]); }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        int p = (int)base + u * 32 + lane;
        if (p < n && ((aw[u] >> lane) & 1u)) finish(c, eval_ter<SMEM>(c, a[u], b[u]), f, F_TER, p);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        int p = (int)base + u * 32 + lane;
        if (p < n && ((aw[u] >> lane) & 1u)) {
          const int4* q = &f.desc[3 * (size_t)p];
          finish(c, eval_dj<SMEM>(c, __ldg(q), __ldg(q + 1), __ldg(q + 2)), f, F_DJ, p);
        }
      }
    }
  }
}

// Tail propagators (allocated after the CSR was built, e.g. the branching constraints of
// search/branching/binary_split.rs:46-57): evaluated every iteration, straight from L2.
__device__ __forceinline__ void eval_tail(Ctx& c) {
  const Params& P = *c.P;
  for (int p = P.bin.n_static + threadIdx.x; p < P.bin.n; p += blockDim.x)
    if (is_active(P.bin, p)) eval_ref<false>(c, F_BIN, p);
  for (int p = P.ter.n_static + threadIdx.x; p < P.ter.n; p += blockDim.x)
    if (is_active(P.ter, p)) eval_ref<false>(c, F_TER, p);
  for (int p = P.dj.n_static + threadIdx.x; p < P.dj.n; p += blockDim.x)
    if (is_active(P.dj, p)) eval_ref<false>(c, F_DJ, p);
}

// ---------------------------------------------------------------------------------------
// n-ary Distinct (propagators/distinct.rs:69-126 = Conjunction of pairwise XNeqY,
// logic/conjunction.rs:77-105).  Fixpoint characterisation: S = values of
// the singleton operands; two equal singletons fail; every other operand advances
// lo while lo in S and retreats hi while hi in S; an operand that becomes a singleton
// joins S.  One CTA per propagator: operands staged in shared memory, S is an
// open-addressing hash set in shared memory, warp ballots decide the rounds.
// ---------------------------------------------------------------------------------------
constexpr int kHashEmpty = INT32_MIN;

__device__ __forceinline__ unsigned hash_slot(int v, unsigned mask) {
  return ((unsigned)v * 2654435761u >> 7) & mask;
}
// returns false if v was already present
__device__ __forceinline__ bool hs_insert(int* tab, unsigned mask, int v) {
  unsigned h = hash_slot(v, mask);
  while (true) {
    int old = atomicCAS(&tab[h], kHashEmpty, v);
    if (old == kHashEmpty) return true;
    if (old == v) return false;
    h = (h + 1) & mask;
  }
}
__device__ __forceinline__ bool hs_contains(const volatile int* tab, unsigned mask, int v) {
  unsigned h = hash_slot(v, mask);
  while (true) {
    int cur = tab[h];
    if (cur == v) return true;
    if (cur == kHashEmpty) return false;
    h = (h + 1) & mask;
  }
}

// smem layout for the n-ary stage (after the optional domain snapshot):
//   int2 ops[k]; int2 iv[k] (view-space lo/hi); int tab[tabsz];
__device__ void eval_distinct(Ctx& c, int slot, char* smem_nary, unsigned cur_epoch, bool first_iter) {
  const Params& P = *c.P;
  const int b = __ldg(&P.nary_ptr[slot]), e = __ldg(&P.nary_ptr[slot + 1]);
  const int k = e - b;
  int2* ops = reinterpret_cast<int2*>(smem_nary);
  int2* iv = ops + P.nary_max_k;
  int* tab = reinterpret_cast<int*>(iv + P.nary_max_k);
  unsigned tabsz = 4;
  while (tabsz < 2u * (unsigned)k) tabsz <<= 1;
  const unsigned mask = tabsz - 1;
  // stage operands + current domains; is any operand dirty this iteration?
  int any_dirty = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    int2 op = __ldg(&P.nary_ops[b + i]);
    ops[i] = op;
    if (op.x >= 0) {
      int2 d = ldcg_dom(&P.dom[op.x]);
      iv[i] = make_int2(d.x + op.y, d.y + op.y);
      if (__ldcg(&P.dirty_stamp[op.x]) == cur_epoch) any_dirty = 1;
    } else {
      iv[i] = make_int2(op.y, op.y);
    }
  }
  for (unsigned i = threadIdx.x; i < tabsz; i += blockDim.x) tab[i] = kHashEmpty;
  if (!first_iter) {
    if (!__syncthreads_or(any_dirty)) return;  // nothing it depends on changed (distinct.rs:119-126)
  } else {
    __syncthreads();
  }
  // rounds: insert new singletons, then prune bounds against S
  // inserted[i] is tracked in the sign of ops[i].x? keep a separate pass marker: iv.x > iv.y never
  // happens for live operands, so we mark "inserted" by remembering it in a register per strided i.
  // (each thread owns indices i = tid, tid+blockDim, ... for the whole evaluation)
  unsigned inserted_bits = 0;  // bit j: operand tid + j*blockDim already in S (k <= 32*blockDim)
  while (true) {
    int fail = 0;
    int j = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x, ++j) {
      int2 d = iv[i];
      if (d.x == d.y && !((inserted_bits >> j) & 1u)) {
        inserted_bits |= 1u << j;
        if (!hs_insert(tab, mask, d.x)) fail = 1;  // two equal singletons: XNeqY fails
      }
    }
    if (__syncthreads_or(fail)) { set_failed(c); if (threadIdx.x == 0 && c.count) c.nprop++; return; }
    int again = 0;
    j = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x, ++j) {
      int2 d = iv[i];
      if (d.x == d.y) continue;
      int lo = d.x, hi = d.y;
      while (lo <= hi && hs_contains(tab, mask, lo)) ++lo;
      while (hi >= lo && hs_contains(tab, mask, hi)) --hi;
      if (lo > hi) { fail = 1; continue; }
      if (lo != d.x || hi != d.y) {
        iv[i] = make_int2(lo, hi);
        if (lo == hi) again = 1;
      }
    }
    // (__syncthreads_or yields a predicate, not the OR of the operands: two reductions)
    if (__syncthreads_or(fail)) { set_failed(c); if (threadIdx.x == 0 && c.count) c.nprop++; return; }
    if (!__syncthreads_or(again)) break;
  }
  // write back narrowed bounds
  long long sum_size = 0;
  int mn = INT32_MAX, mx = INT32_MIN;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    int2 op = ops[i];
    int2 d = iv[i];
    if (op.x >= 0) {
      int2 cur = ldcg_dom(&P.dom[op.x]);
      IV curv{cur.x + op.y, cur.y + op.y};
      if (!tighten(c, op.x, op.y, curv, max(curv.lo, d.x), min(curv.hi, d.y))) set_failed(c);
    }
    sum_size += (long long)d.y - d.x + 1;
    mn = min(mn, d.x);
    mx = max(mx, d.y);
  }
  // entailment: Conjunction::is_subsumed (conjunction.rs:77-94) = all pairs disjoint.
  // Necessary condition first (pigeonhole): sum of sizes <= span.
  __shared__ long long s_sum;
  __shared__ int s_mn, s_mx;
  if (threadIdx.x == 0) { s_sum = 0; s_mn = INT32_MAX; s_mx = INT32_MIN; }
  __syncthreads();
  for (int o = 16; o; o >>= 1) {
    sum_size += __shfl_xor_sync(0xffffffffu, sum_size, o);
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd((unsigned long long*)&s_sum, (unsigned long long)sum_size);
    atomicMin(&s_mn, mn);
    atomicMax(&s_mx, mx);
  }
  __syncthreads();
  bool entailed = false;
  if (k <= 1) {
    entailed = true;  // empty conjunction (distinct_test case 7)
  } else if (s_sum <= (long long)s_mx - s_mn + 1) {
    int overlap = 0;
    for (int i = threadIdx.x; i < k && !overlap; i += blockDim.x) {
      int2 a = iv[i];
      for (int q = i + 1; q < k; ++q) {
        int2 bq = iv[q];
        if (!(a.y < bq.x || bq.y < a.x)) { overlap = 1; break; }
      }
    }
    entailed = !__syncthreads_or(overlap);
  }
  if (threadIdx.x == 0 && c.count) {
    c.nprop++;
    if (entailed) {
      unsigned bit = 1u << (slot & 31);
      unsigned old = atomicAnd(&P.nary_active[slot >> 5], ~bit);
      if (old & bit) P.trail[atomicAdd(&P.ctl->trail_cnt, 1u)] = make_ref(F_NARY, (unsigned)slot);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// device-wide barrier; the last CTA to arrive decides whether the fixpoint is reached
// (the "block-reduce of a changed flag": the reduction operand is the dirty-list length).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned barrier_decide(const Params& P, unsigned& gen, unsigned* s_block_props,
                                                  int next_buf, unsigned iter) {
  __shared__ unsigned s_dec;
  __syncthreads();
  if (threadIdx.x == 0) {
    Control* ctl = P.ctl;
    if (*s_block_props) { atomicAdd(&ctl->propagations, (unsigned long long)*s_block_props); *s_block_props = 0; }
    __threadfence();
    unsigned arrived = atomicAdd(&ctl->bar_count, 1u);
    if (arrived == gridDim.x - 1) {
      unsigned dec;
      int failed = *(volatile int*)&ctl->failed;
      int nd = *(volatile int*)&ctl->dirty_cnt[next_buf];
      if (failed) dec = D_FAILED;
      else if (nd == 0) dec = D_FIXPOINT;
      else if (iter + 1 >= P.max_iterations) dec = D_ITER_CAP;
      else dec = D_CONTINUE;
      ctl->decision[gen & 1u] = dec;
      ctl->bar_count = 0;
      __threadfence();
      atomicAdd(&ctl->bar_gen, 1u);
      s_dec = dec;
    } else {
      while (ld_acquire_u32(&ctl->bar_gen) == gen) { }
      s_dec = *(volatile unsigned*)&ctl->decision[gen & 1u];
    }
    __threadfence();
  }
  __syncthreads();
  gen++;
  return s_dec;
}

template <bool SMEM>
__device__ __forceinline__ void expand_dirty_rows(Ctx& c, int cur_buf, int n_dirty, unsigned cur_epoch) {
  const Params& P = *c.P;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarps;
  const long long S = max(1LL, nwarps / n_dirty);  // segments per row (upper bound)
  const long long items = (long long)n_dirty * S;
  const int* list = P.dirty_list + (size_t)cur_buf * P.V;
  for (long long item = warp; item < items; item += nwarps) {
    int e = (int)(item / S);
    long long s = item % S;
    int v = __ldcg(&list[e]);
    int rb = __ldg(&P.adj_ptr[v]), re = __ldg(&P.adj_ptr[v + 1]);
    long long len = re - rb;
    long long nseg = min(S, (len + 31) / 32);
    if (s >= nseg) continue;
    int sb = rb + (int)(len * s / nseg), se = rb + (int)(len * (s + 1) / nseg);
    for (int j = sb + lane; j < se; j += 32) {
      unsigned ref = __ldg(&P.adj[j]);
      unsigned fam = ref >> 29;
      int slot = (int)(ref & kSlotMask);
      const Family& f = fam == F_BIN ? P.bin : (fam == F_TER ? P.ter : P.dj);
      if (slot >= f.n_static) continue;  // truncated by a restore (store.rs:320)
      if (!is_active(f, slot)) continue;
      if (atomicExch(&f.stamp[slot], cur_epoch) == cur_epoch) continue;  // already scheduled
      eval_ref<SMEM>(c, fam, slot);
    }
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(kThreads, 1) pcp_fixpoint_kernel(const Params P) {
  extern __shared__ __align__(16) char smem[];
  __shared__ unsigned s_block_props;
  int2* sdom = SMEM ? reinterpret_cast<int2*>(smem) : nullptr;
  char* smem_nary = smem + (SMEM ? (size_t)((P.V * 8 + 15) & ~15) : 0);

  Control* ctl = P.ctl;
  unsigned gen = 0;
  unsigned epoch0 = 0;
  if (threadIdx.x == 0) {
    s_block_props = 0;
    gen = *(volatile unsigned*)&ctl->bar_gen;
    epoch0 = *(volatile unsigned*)&ctl->epoch;
  }
  {
    __shared__ unsigned s_gen, s_epoch;
    if (threadIdx.x == 0) { s_gen = gen; s_epoch = epoch0; }
    __syncthreads();
    gen = s_gen;
    epoch0 = s_epoch;
  }

  Ctx c;
  c.P = &P;
  c.sdom = sdom;
  c.nprop = 0;

  unsigned iter = 0;
  unsigned dec;
  const int lane = threadIdx.x & 31;
  while (true) {
    const int cur_buf = iter % 3, next_buf = (iter + 1) % 3, spare_buf = (iter + 2) % 3;
    const unsigned cur_epoch = epoch0 + iter;
    c.next_epoch = cur_epoch + 1;
    c.next_buf = next_buf;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->dirty_cnt[spare_buf] = 0;  // idle this iteration

    // 1. tail propagators: iteration 0 on every CTA (so each snapshot already holds the
    //    branching decision), afterwards CTA 0 only.  Counted once.
    if (iter == 0 || blockIdx.x == 0) {
      c.count = blockIdx.x == 0;
      eval_tail(c);
      c.count = true;
      if (iter == 0) { __threadfence(); __syncthreads(); }
    }

    if (iter == 0) {
      if (SMEM) {
        for (int v = threadIdx.x; v < P.V; v += blockDim.x) sdom[v] = ldcg_dom(&P.dom[v]);
        __syncthreads();
      }
      if (P.full_sweep) {
        sweep_family<SMEM, F_BIN>(c, P.bin);
        sweep_family<SMEM, F_TER>(c, P.ter);
        sweep_family<SMEM, F_DJ>(c, P.dj);
      }
    }
    // worklist of variables narrowed in the previous iteration (iteration 0 of an
    // incremental launch: seeded by the host)
    int n_dirty = 0;
    if (iter > 0 || !P.full_sweep) n_dirty = *(volatile int*)&ctl->dirty_cnt[cur_buf];
    if (n_dirty > 0) {
      const int* list = P.dirty_list + (size_t)cur_buf * P.V;
      // refresh the snapshot and catch domains emptied by two concurrent updates
      int bad = 0;
      for (int i = threadIdx.x; i < n_dirty; i += blockDim.x) {
        int v = __ldcg(&list[i]);
        int2 d = ldcg_dom(&P.dom[v]);
        if (SMEM) sdom[v] = d;
        bad |= d.x > d.y;
      }
      if (bad) set_failed(c);
      if (SMEM) __syncthreads();
      expand_dirty_rows<SMEM>(c, cur_buf, n_dirty, cur_epoch);
    }
    // 2. n-ary propagators: one CTA each; re-run when one of their operands is dirty.
    if (P.n_nary > 0 && (iter > 0 || P.full_sweep || n_dirty > 0)) {
      // fold this thread's count first (eval_distinct counts on thread 0)
      for (int s = blockIdx.x; s < P.n_nary; s += gridDim.x) {
        if (!((__ldcg(&P.nary_active[s >> 5]) >> (s & 31)) & 1u)) continue;
        eval_distinct(c, s, smem_nary, cur_epoch, iter == 0 && P.full_sweep);
      }
    }

    // block-level propagation count, then the barrier + decision
    unsigned np = c.nprop;
    c.nprop = 0;
    for (int o = 16; o; o >>= 1) np += __shfl_xor_sync(0xffffffffu, np, o);
    if (lane == 0 && np) atomicAdd(&s_block_props, np);
    dec = barrier_decide(P, gen, &s_block_props, next_buf, iter);
    ++iter;
    if (dec != D_CONTINUE) break;
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->epoch = epoch0 + iter + 1;
    ctl->iterations = iter;
    ctl->last_decision = dec;
    ctl->dirty_cnt[0] = ctl->dirty_cnt[1] = ctl->dirty_cnt[2] = 0;
    Result r;
    r.failed = dec == D_FAILED;
    r.trail_cnt = *(volatile unsigned*)&ctl->trail_cnt;
    r.iterations = iter;
    r.epoch = epoch0 + iter + 1;
    r.propagations = *(volatile unsigned long long*)&ctl->propagations;
    r.decision = dec;
    *P.result = r;
  }
}

// ---------------------------------------------------------------------------------------
// node prologue: Snapshot::restore (+ Store::alloc of the few propagators posted since),
// in one launch: domains <- label copy, `active` bits of the trail suffix set again
// (propagation/store.rs:319-323), active bits of newly allocated tail slots set.
// ---------------------------------------------------------------------------------------
struct NodeBegin {
  int2* dom;
  const int2* restore_from;  // nullptr: keep the current domains
  int V;
  uint32_t* trail;
  unsigned trail_keep;       // trail length recorded in the label
  int do_trail;              // 1: undo trail entries >= trail_keep
  uint32_t* active[4];       // per family (BIN, TER, DJ, NARY)
  int new_first[4], new_last[4];  // slots whose active bit must be set
  Control* ctl;
};

__global__ void __launch_bounds__(1024, 1) pcp_node_begin_kernel(const NodeBegin nb) {
  const int tid = threadIdx.x, nth = blockDim.x;  // single CTA
  if (nb.restore_from)
    for (int v = tid; v < nb.V; v += nth) nb.dom[v] = nb.restore_from[v];
  if (nb.do_trail) {
    unsigned cnt = nb.ctl->trail_cnt;
    for (unsigned i = nb.trail_keep + tid; i < cnt; i += nth) {
      unsigned ref = nb.trail[i];
      unsigned slot = ref & kSlotMask;
      atomicOr(&nb.active[ref >> 29][slot >> 5], 1u << (slot & 31));
    }
  }
  for (int f = 0; f < 4; ++f) {
    const int first = nb.new_first[f], last = nb.new_last[f];
    if (first >= last) continue;
    for (int w = (first >> 5) + tid; w <= ((last - 1) >> 5); w += nth) {
      int lo = max(first, w * 32) - w * 32, hi = min(last, w * 32 + 32) - w * 32;  // bits [lo, hi)
      unsigned mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
      atomicOr(&nb.active[f][w], mask);
    }
  }
  __syncthreads();
  if (tid == 0) {
    nb.ctl->failed = 0;
    if (nb.do_trail) nb.ctl->trail_cnt = nb.trail_keep;
    nb.ctl->dirty_cnt[0] = nb.ctl->dirty_cnt[1] = nb.ctl->dirty_cnt[2] = 0;
  }
}

__global__ void pcp_fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace pcpd
