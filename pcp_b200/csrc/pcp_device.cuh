// pcp_device.cuh -- sm_100a device side of the propagation engine.
//
// One persistent kernel (`pcp_fixpoint_kernel`) runs the whole propagation fixpoint of a
// search node -- the device equivalent of Store::consistency
// (reference src/libpcp/propagation/store.rs:247-257): one CTA per SM, a device-wide
// barrier between iterations, quiescence detected by the last CTA to arrive.
//
//   prologue      Snapshot::restore + Store::alloc of the few propagators posted since the
//                 last call (CTA 0), overlapped with the first descriptor loads.
//   iteration 0   every active propagator is evaluated once (store.rs:144-149 schedules
//                 all active propagators).  The per-family descriptor arrays are streamed
//                 from HBM by TMA bulk copies (cp.async.bulk + mbarrier, 4-stage ring in
//                 shared memory, one producer warp, 31 consumer warps); domains are read
//                 from a shared-memory snapshot (V <= ~12k) or gathered from L2.
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
constexpr int kConsumerWarps = kWarps - 1;  // warp 0 drives the TMA ring
constexpr int kStages = 4;
constexpr int kStageBytes = 32768;
constexpr int kRingBytes = kStages * kStageBytes;
// propagators per ring stage: multiples of 32 * kConsumerWarps so every consumer warp gets
// whole 32-propagator groups (aligned with the words of the `active` bit set)
constexpr int kChunkBin = 1984;         // 1984 * 16 B = 31744 B
constexpr int kChunkTer = 992;          // 992 * 16 B + 992 * 8 B
constexpr int kChunkDj = 480;           // 480 * 48 B = 23040 B
constexpr int kTerPlaneB = 16384;       // offset of the z plane inside a stage
constexpr unsigned kConstVar28 = 0x0FFFFFFFu;
constexpr unsigned kSumBase28 = 0x0F000000u;  // x-operand encodings >= this are sum views / Constant
constexpr int kMaxInline = 4;

// family tags inside adjacency / trail references (top 3 bits)
enum Fam : unsigned { F_BIN = 0, F_TER = 1, F_DJ = 2, F_NARY = 3 };
constexpr unsigned kSlotMask = 0x1FFFFFFFu;
__host__ __device__ inline unsigned make_ref(unsigned fam, unsigned slot) { return (fam << 29) | slot; }

// kinds inside a family (top 4 bits of descriptor word 0)
enum BinKind : unsigned { B_LESS = 0, B_NEQ = 1, B_EQ = 2 };
enum TerKind : unsigned { T_GREATER = 0, T_LESS = 1, T_EQ = 2 };

enum Decision : unsigned { D_CONTINUE = 0, D_FIXPOINT = 1, D_FAILED = 2, D_ITER_CAP = 3, D_SWEEP = 4 };
constexpr unsigned kDecBits = 3;  // the release word of the barrier = (generation << kDecBits) | decision

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
  int4* desc;           // BIN: 1 int4/prop; TER: int4 (x,y) plane; DJ: 3 int4/prop
  int2* descB;          // TER: (z) plane
  uint32_t* active;     // bit set, 1 = active (propagation/store.rs:34)
  uint32_t* stamp;      // epoch of the last worklist evaluation
  int n;                // allocated propagators
  int n_static;         // [0, n_static) are covered by the CSR; [n_static, n) is the tail
};

struct InlineProp {     // a propagator posted since the last launch, carried in the launch
  int4 q[3];            // parameters instead of a separate H2D copy
  unsigned fam;
  int slot;
  int pad[2];
};

struct Params {
  Result* result;
  int2* dom;            // interval per variable: (lo, hi)
  int V;
  int smem_dom;         // 1: domains are staged in shared memory
  Family fam[3];        // BIN, TER, DJ
  const int* nary_ptr;  // CSR of n-ary Distinct operands
  const int2* nary_ops;
  uint32_t* nary_active;
  int n_nary;
  int nary_max_k;
  const int* sum_ptr;   // CSR of the Sum views (term/sum.rs): terms are (var, off) operands
  const int2* sum_terms;
  const int* adj_ptr;   // reactor: var -> propagator refs
  const uint32_t* adj;
  int* dirty_list;      // 3 x V
  uint32_t* dirty_stamp;
  uint32_t* trail;
  Control* ctl;
  int full_sweep;       // 1: schedule every active propagator first (store.rs:144-149)
  unsigned max_iterations;
  // ---- node prologue (Snapshot::restore + Store::alloc), executed by CTA 0
  const int2* restore_from;   // label copy of the domains, or nullptr
  unsigned trail_keep;        // trail length recorded in the label
  int do_trail;               // 1: re-activate trail entries >= trail_keep (store.rs:319-323)
  int sync0;                  // 1: the prologue has cross-CTA effects -> barrier before use
  uint32_t* nary_active_w;    // == nary_active (writable alias for the prologue)
  int new_first[4], new_last[4];  // slots whose active bit must be set (per family)
  int n_inline;
  InlineProp inl[kMaxInline];
  int seed_dirty;             // incremental launch: dirty_list[0..seed_dirty) seeded by the host
  // ---- epilogue
  int2* snapshot_to;          // if not failed: copy of the fixpoint domains (next label slot)
  // ---- debug: per-CTA phase timestamps (PCP_TRACE=1), 8 slots per CTA
  unsigned long long* trace;
};

__device__ __forceinline__ void trace_mark(const Params& P, int slot) {
  if (P.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[blockIdx.x * 8 + slot] = t;
  }
}

// ---------------------------------------------------------------------------------------
// memory / async helpers
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
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace pcpd

#include "pcp_eval.cuh"

namespace pcpd {

// ---------------------------------------------------------------------------------------
// iteration 0: the streaming sweep.  A chunk = up to kChunk* propagators of one family;
// chunk g belongs to CTA g % gridDim.x.  Warp 0 (one lane) is the producer: it arms the
// stage's `full` mbarrier with the byte count and issues the TMA bulk copies; the 31
// consumer warps wait on `full`, evaluate from shared memory and release the stage through
// the `empty` mbarrier.
// ---------------------------------------------------------------------------------------
struct ChunkMap {
  int nch[3];   // chunks per family
  int total;
};
__device__ __forceinline__ ChunkMap chunk_map(const Params& P) {
  ChunkMap m;
  m.nch[0] = (P.fam[0].n_static + kChunkBin - 1) / kChunkBin;
  m.nch[1] = (P.fam[1].n_static + kChunkTer - 1) / kChunkTer;
  m.nch[2] = (P.fam[2].n_static + kChunkDj - 1) / kChunkDj;
  m.total = m.nch[0] + m.nch[1] + m.nch[2];
  return m;
}
struct Chunk { int fam, base, cnt; };
__device__ __forceinline__ Chunk chunk_of(const Params& P, const ChunkMap& m, int g) {
  Chunk c;
  if (g < m.nch[0]) { c.fam = 0; c.base = g * kChunkBin; c.cnt = min(kChunkBin, P.fam[0].n_static - c.base); }
  else if (g < m.nch[0] + m.nch[1]) { g -= m.nch[0]; c.fam = 1; c.base = g * kChunkTer; c.cnt = min(kChunkTer, P.fam[1].n_static - c.base); }
  else { g -= m.nch[0] + m.nch[1]; c.fam = 2; c.base = g * kChunkDj; c.cnt = min(kChunkDj, P.fam[2].n_static - c.base); }
  return c;
}

__device__ __forceinline__ void producer_issue(const Params& P, const Chunk& ch, char* stage, uint64_t* full) {
  const Family& f = P.fam[ch.fam];
  if (ch.fam == F_BIN) {
    unsigned bytes = (unsigned)ch.cnt * 16u;
    mbar_expect_tx(full, bytes);
    bulk_g2s(stage, f.desc + ch.base, bytes, full);
  } else if (ch.fam == F_TER) {
    unsigned ba = (unsigned)ch.cnt * 16u, bb = ((unsigned)ch.cnt * 8u + 15u) & ~15u;
    mbar_expect_tx(full, ba + bb);
    bulk_g2s(stage, f.desc + ch.base, ba, full);
    bulk_g2s(stage + kTerPlaneB, f.descB + ch.base, bb, full);
  } else {
    unsigned bytes = (unsigned)ch.cnt * 48u;
    mbar_expect_tx(full, bytes);
    bulk_g2s(stage, f.desc + 3 * (size_t)ch.base, bytes, full);
  }
}

// The words of the `active` bit set a consumer warp needs for one chunk (at most
// kGroupsMax groups of 32 propagators); loaded one chunk ahead so that the L2 round trip is
// off the critical path of the sweep.
constexpr int kGroupsMax = 2;
struct ActiveWords { unsigned w[kGroupsMax]; };
__device__ __forceinline__ ActiveWords load_active_words(const Params& P, const Chunk& ch) {
  const int cw = (threadIdx.x >> 5) - 1;
  ActiveWords a;
#pragma unroll
  for (int g = 0; g < kGroupsMax; ++g) {
    const int j0 = cw * 32 + g * kConsumerWarps * 32;
    a.w[g] = j0 < ch.cnt ? __ldcg(&P.fam[ch.fam].active[(ch.base + j0) >> 5]) : 0u;
  }
  return a;
}

template <bool SMEM>
__device__ __forceinline__ unsigned sweep_consume(const Ctx& c, const Chunk& ch, const char* stage,
                                                  const ActiveWords& aw) {
  const int lane = threadIdx.x & 31;
  const int cw = (threadIdx.x >> 5) - 1;  // consumer warp index 0..30
  unsigned nprop = 0;
#pragma unroll
  for (int g = 0; g < kGroupsMax; ++g) {
    const int j0 = cw * 32 + g * kConsumerWarps * 32;
    if (j0 >= ch.cnt) break;
    const int j = j0 + lane;
    const unsigned word = aw.w[g];
    if (j >= ch.cnt || !((word >> lane) & 1u)) continue;
    const int slot = ch.base + j;
    ++nprop;
    if (ch.fam == F_BIN) {
      int4 d = reinterpret_cast<const int4*>(stage)[j];
      unsigned kind = (unsigned)d.x >> 28;
      IV x = rd<SMEM>(c, dec_var28((unsigned)d.x), d.y), y = rd<SMEM>(c, d.z, d.w);
      if (!bin_is_noop(kind, x, y)) eval_full<SMEM>(c, F_BIN, slot, d, d, d);
    } else if (ch.fam == F_TER) {
      int4 a = reinterpret_cast<const int4*>(stage)[j];
      int2 b = reinterpret_cast<const int2*>(stage + kTerPlaneB)[j];
      unsigned kind = (unsigned)a.x >> 28;
      IV x = rd<SMEM>(c, dec_var28((unsigned)a.x), a.y), y = rd<SMEM>(c, a.z, a.w), z = rd<SMEM>(c, b.x, b.y);
      if (!ter_is_noop(kind, x, y, z)) eval_full<SMEM>(c, F_TER, slot, a, make_int4(b.x, b.y, 0, 0), a);
    } else {
      const int4* q = reinterpret_cast<const int4*>(stage) + 3 * j;
      int4 q0 = q[0], q1 = q[1], q2 = q[2];
      if (!dj_is_noop<SMEM>(c, q0, q1, q2)) eval_full<SMEM>(c, F_DJ, slot, q0, q1, q2);
    }
  }
  return nprop;
}

// ---------------------------------------------------------------------------------------
// n-ary Distinct (propagators/distinct.rs:69-126 = Conjunction of pairwise XNeqY,
// logic/conjunction.rs:77-105).  Fixpoint characterisation: S = values of
// the singleton operands; two equal singletons fail; every other operand advances
// lo while lo in S and retreats hi while hi in S; an operand that becomes a singleton
// joins S.  One CTA per propagator: operands staged in shared memory, S is an
// open-addressing hash set in shared memory, block-wide votes decide the rounds.
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

// smem layout of the n-ary stage (aliases the TMA ring, idle by then):
//   int2 ops[k]; int2 iv[k] (view-space lo/hi); int tab[tabsz];
// Returns 1 if the propagator was evaluated (thread 0 only), else 0.
template <bool SMEM>
__device__ __noinline__ unsigned eval_distinct(const Ctx& c, int slot, char* smem_nary, unsigned cur_epoch,
                                               bool unconditional) {
  const Params& P = *c.P;
  const int b = __ldg(&P.nary_ptr[slot]), e = __ldg(&P.nary_ptr[slot + 1]);
  const int k = e - b;
  int2* ops = reinterpret_cast<int2*>(smem_nary);
  int2* iv = ops + P.nary_max_k;
  int* tab = reinterpret_cast<int*>(iv + P.nary_max_k);
  unsigned tabsz = 4;
  while (tabsz < 2u * (unsigned)k) tabsz <<= 1;
  const unsigned mask = tabsz - 1;
  __syncthreads();  // previous users of the staging area are done
  // stage operands + current domains; is any operand dirty this iteration?
  int any_dirty = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    int2 op = __ldg(&P.nary_ops[b + i]);
    ops[i] = op;
    if (op.x >= 0) {
      int2 d = SMEM ? c.sdom[op.x] : ldcg_dom(&P.dom[op.x]);
      iv[i] = make_int2(d.x + op.y, d.y + op.y);
      if (__ldcg(&P.dirty_stamp[op.x]) == cur_epoch) any_dirty = 1;
    } else {
      iv[i] = make_int2(op.y, op.y);
    }
  }
  for (unsigned i = threadIdx.x; i < tabsz; i += blockDim.x) tab[i] = kHashEmpty;
  if (!unconditional) {
    if (!__syncthreads_or(any_dirty)) return 0;  // nothing it depends on changed (distinct.rs:119-126)
  } else {
    __syncthreads();
  }
  // rounds: insert new singletons, then prune bounds against S.  Each thread owns the operands
  // i = tid, tid + blockDim, ... for the whole evaluation (k <= 32 * blockDim).
  unsigned inserted_bits = 0;
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
    if (__syncthreads_or(fail)) { set_failed(c); return 1; }
    int again = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
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
    if (__syncthreads_or(fail)) { set_failed(c); return 1; }
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
  if (threadIdx.x == 0 && entailed) {
    unsigned bit = 1u << (slot & 31);
    unsigned old = atomicAnd(&P.nary_active[slot >> 5], ~bit);
    if (old & bit) P.trail[atomicAdd(&P.ctl->trail_cnt, 1u)] = make_ref(F_NARY, (unsigned)slot);
  }
  return 1;
}

// ---------------------------------------------------------------------------------------
// device-wide barrier; the last CTA to arrive decides whether the fixpoint is reached
// (the "block-reduce of a changed flag": the reduction operand is the dirty-list length).
// `decide` = false: plain barrier (after the node prologue).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void node_prologue_finish(const Params& P);
// The arrival word packs three 10-bit counters -- CTAs arrived, CTAs that queued a dirty
// variable, CTAs that saw a failure -- so one release-atomic per CTA carries everything the
// decision needs; the release word `bar_gen` = (generation << kDecBits) | decision, so one
// acquire-load per poll returns both.
__device__ __forceinline__ unsigned grid_barrier(const Params& P, unsigned& gen, unsigned block_props, bool decide,
                                                 const int* s_flags, unsigned iter, int next_buf = 0) {
  __shared__ unsigned s_dec;
  __syncthreads();
  if (threadIdx.x == 0) {
    Control* ctl = P.ctl;
    if (block_props) atomicAdd(&ctl->propagations, (unsigned long long)block_props);
    unsigned add = 1u + (s_flags[0] ? (1u << 10) : 0u) + (s_flags[1] ? (1u << 20) : 0u);
    const_cast<int*>(s_flags)[0] = const_cast<int*>(s_flags)[1] = 0;  // nobody sets them inside the barrier
    unsigned old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(&ctl->bar_count), "r"(add) : "memory");
    unsigned now = old + add;
    if ((now & 1023u) == gridDim.x) {
      unsigned dec = D_CONTINUE;
      if (decide) {
        if ((now >> 20) & 1023u) dec = D_FAILED;
        else if (((now >> 10) & 1023u) == 0) dec = D_FIXPOINT;
        else if (iter + 1 >= P.max_iterations) dec = D_ITER_CAP;
        else {
          // many dirty variables: their CSR rows cover most of the store, and a second streaming
          // sweep is cheaper than gathering the rows
          int nd = *(volatile int*)&ctl->dirty_cnt[next_buf];
          if ((long long)nd * 8 >= (long long)P.V) dec = D_SWEEP;
        }
      }
      if (!decide) node_prologue_finish(P);  // prologue barrier: every CTA has read trail_cnt by now
      ctl->bar_count = 0;
      unsigned rel = ((gen + 1u) << kDecBits) | dec;
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&ctl->bar_gen), "r"(rel) : "memory");
      s_dec = dec;
    } else {
      unsigned v;
      do { v = ld_acquire_u32(&ctl->bar_gen); } while ((v >> kDecBits) == gen);
      s_dec = v & ((1u << kDecBits) - 1u);
    }
  }
  __syncthreads();
  gen = (gen + 1u) & (0xffffffffu >> kDecBits);
  return s_dec;
}

template <bool SMEM>
__device__ __forceinline__ unsigned expand_dirty_rows(const Ctx& c, int cur_buf, int n_dirty, unsigned cur_epoch) {
  const Params& P = *c.P;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarps;
  const long long S = max(1LL, nwarps / n_dirty);  // segments per row (upper bound)
  const long long items = (long long)n_dirty * S;
  const int* list = P.dirty_list + (size_t)cur_buf * P.V;
  unsigned nprop = 0;
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
      const Family& f = P.fam[fam];
      if (slot >= f.n_static) continue;  // truncated by a restore (store.rs:320)
      // the active word and the descriptor are independent loads: both in flight together
      const unsigned word = __ldcg(&f.active[slot >> 5]);
      int4 q0, q1, q2;
      load_desc(f, fam, slot, q0, q1, q2);
      if (!((word >> (slot & 31)) & 1u)) continue;
      if (atomicExch(&f.stamp[slot], cur_epoch) == cur_epoch) continue;  // already scheduled
      eval_loaded<SMEM>(c, fam, slot, q0, q1, q2);
      ++nprop;
    }
  }
  return nprop;
}

// Node prologue: Snapshot::restore (domains <- label copy, `active` bits of the trail suffix
// set again, propagation/store.rs:319-323) and Store::alloc of the propagators posted since
// the last launch (descriptor + active bit).  `tid`/`nth` span CTA 0 when nothing crosses
// CTAs, or the whole grid when a device barrier follows anyway (a restore over a long trail
// -- e.g. a store that had become fully entailed -- is then a grid-wide job).  The caller
// resets `trail_cnt` once every participant has read it (node_prologue_finish).
__device__ __forceinline__ void node_prologue(const Params& P, int tid, int nth) {
  if (P.restore_from)
    for (int v = tid; v < P.V; v += nth) P.dom[v] = P.restore_from[v];
  if (P.do_trail) {
    unsigned cnt = *(volatile unsigned*)&P.ctl->trail_cnt;
    for (unsigned i = P.trail_keep + tid; i < cnt; i += nth) {
      unsigned ref = P.trail[i];
      unsigned fam = ref >> 29, slot = ref & kSlotMask;
      uint32_t* act = fam == F_NARY ? P.nary_active_w : P.fam[fam].active;
      atomicOr(&act[slot >> 5], 1u << (slot & 31));
    }
  }
  for (int f = 0; f < 4; ++f) {
    const int first = P.new_first[f], last = P.new_last[f];
    if (first >= last) continue;
    uint32_t* act = f == F_NARY ? P.nary_active_w : P.fam[f].active;
    for (int w = (first >> 5) + tid; w <= ((last - 1) >> 5); w += nth) {
      int lo = max(first, w * 32) - w * 32, hi = min(last, w * 32 + 32) - w * 32;  // bits [lo, hi)
      unsigned mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
      atomicOr(&act[w], mask);
    }
  }
  if (tid < P.n_inline) {
    const InlineProp& ip = P.inl[tid];
    const Family& f = P.fam[ip.fam];
    if (ip.fam == F_BIN) f.desc[ip.slot] = ip.q[0];
    else if (ip.fam == F_TER) { f.desc[ip.slot] = ip.q[0]; f.descB[ip.slot] = make_int2(ip.q[1].x, ip.q[1].y); }
    else { f.desc[3 * (size_t)ip.slot] = ip.q[0]; f.desc[3 * (size_t)ip.slot + 1] = ip.q[1]; f.desc[3 * (size_t)ip.slot + 2] = ip.q[2]; }
  }
}
__device__ __forceinline__ void node_prologue_finish(const Params& P) {
  if (P.do_trail) P.ctl->trail_cnt = P.trail_keep;
}

template <bool SMEM>
__global__ void __launch_bounds__(kThreads, 1) pcp_fixpoint_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
  __shared__ unsigned s_block_props, s_gen, s_epoch;
  __shared__ int s_flags[2];
  // layout: [ring | n-ary staging (aliased)] [domain snapshot]
  char* ring = smem;
  int2* sdom = SMEM ? reinterpret_cast<int2*>(smem + kRingBytes) : nullptr;

  Control* ctl = P.ctl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // CTA 0 keeps the books (prologue, posted + tail propagators, result); the sweep is shared
  // by the other CTAs so that nobody waits for it at the barrier
  const ChunkMap cmap = chunk_map(P);
  const int workers = gridDim.x > 1 ? (int)gridDim.x - 1 : 1;
  const int wid = gridDim.x > 1 ? (int)blockIdx.x - 1 : 0;
  const int my_chunks = wid >= 0 && wid < cmap.total ? (cmap.total - 1 - wid) / workers + 1 : 0;
  const int pre_issued = P.full_sweep ? min(my_chunks, kStages) : 0;
  int pipe_pos = 0;  // chunks this CTA has pushed through the ring so far (all sweeps)

  if (threadIdx.x == 0) {
    s_block_props = 0;
    s_flags[0] = s_flags[1] = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // start streaming descriptors right away: they do not depend on the node prologue
    for (int i = 0; i < pre_issued; ++i)
      producer_issue(P, chunk_of(P, cmap, wid + i * workers), ring + i * kStageBytes, &s_full[i]);
  } else if (threadIdx.x == 32) {
    s_gen = *(volatile unsigned*)&ctl->bar_gen >> kDecBits;
    s_epoch = *(volatile unsigned*)&ctl->epoch;
  }
  // without a restore the domains are already final: snapshot them while the TMA runs
  if (SMEM && !P.sync0)
    for (int v = threadIdx.x; v < P.V; v += blockDim.x) sdom[v] = ldcg_dom(&P.dom[v]);
  __syncthreads();
  unsigned gen = s_gen;
  const unsigned epoch0 = s_epoch;
  trace_mark(P, 0);

  if (P.sync0) {
    node_prologue(P, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    grid_barrier(P, gen, 0, false, s_flags, 0);  // its last arriver resets trail_cnt (nobody pushes yet)
    if (SMEM) {
      for (int v = threadIdx.x; v < P.V; v += blockDim.x) sdom[v] = ldcg_dom(&P.dom[v]);
      __syncthreads();
    }
  } else if (blockIdx.x == 0) {
    node_prologue(P, threadIdx.x, blockDim.x);  // CTA-local effects only (posted propagators)
    __syncthreads();
  }
  trace_mark(P, 1);
  ActiveWords aw;
  aw.w[0] = aw.w[1] = 0u;
  if (warp > 0 && my_chunks > 0 && P.full_sweep) aw = load_active_words(P, chunk_of(P, cmap, wid));
  bool sweep_now = P.full_sweep != 0;

  Ctx c;
  c.P = &P;
  c.sdom = sdom;
  c.flags = s_flags;

  unsigned iter = 0;
  unsigned dec;
  unsigned nprop = 0;
  while (true) {
    const int cur_buf = iter % 3, next_buf = (iter + 1) % 3, spare_buf = (iter + 2) % 3;
    const unsigned cur_epoch = epoch0 + iter;
    c.next_epoch = cur_epoch + 1;
    c.next_buf = next_buf;
    c.local = false;
    c.mark_dirty = true;
    c.bookkeep = true;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->dirty_cnt[spare_buf] = 0;  // idle this iteration

    if (iter == 0 && P.n_inline > 0) {
      // Propagators posted since the last launch.  With a full sweep ahead they need not enter
      // the worklist: every CTA applies them to its own view of the domains first, so the sweep
      // already sees their effect.
      if (threadIdx.x == 0) {
        c.mark_dirty = !P.full_sweep;
        if (blockIdx.x == 0 || !SMEM) {  // the global store: counted and book-kept by CTA 0
          c.bookkeep = blockIdx.x == 0;
          for (int i = 0; i < P.n_inline; ++i)
            eval_full<false>(c, P.inl[i].fam, P.inl[i].slot, P.inl[i].q[0], P.inl[i].q[1], P.inl[i].q[2]);
          if (blockIdx.x == 0) nprop += P.n_inline;
          if (!SMEM) __threadfence();
        }
        if (SMEM) {
          c.local = true;
          c.bookkeep = false;
          for (int i = 0; i < P.n_inline; ++i)
            eval_full<true>(c, P.inl[i].fam, P.inl[i].slot, P.inl[i].q[0], P.inl[i].q[1], P.inl[i].q[2]);
        }
        c.local = false;
        c.mark_dirty = true;
        c.bookkeep = true;
      }
      __syncthreads();
    }
    // older tail propagators: CTA 0, every iteration (inline ones were just handled)
    if (blockIdx.x == 0) {
      unsigned n = 0;
      for (unsigned fam = 0; fam < 3; ++fam) {
        const Family& f = P.fam[fam];
        for (int p = f.n_static + threadIdx.x; p < f.n; p += blockDim.x) {
          bool inl = false;
          if (iter == 0)
            for (int i = 0; i < P.n_inline; ++i) inl |= P.inl[i].fam == fam && P.inl[i].slot == p;
          if (!inl && is_active(f, p)) { eval_ref<false>(c, fam, p); ++n; }
        }
      }
      nprop += n;
    }

    if (iter == 0) trace_mark(P, 2);
    // worklist of variables narrowed in the previous iteration (iteration 0 of an
    // incremental launch: seeded by the host)
    int n_dirty = 0;
    if (iter > 0) n_dirty = *(volatile int*)&ctl->dirty_cnt[cur_buf];
    else if (!P.full_sweep) n_dirty = P.seed_dirty;
    if (n_dirty > 0) {
      const int* list = P.dirty_list + (size_t)cur_buf * P.V;
      // refresh the snapshot and catch domains emptied by two concurrent updates
      // (with a snapshot every CTA needs every dirty variable; without one the check is
      // shared by the grid)
      int bad = 0;
      const int r0 = SMEM ? threadIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
      const int rs = SMEM ? blockDim.x : gridDim.x * blockDim.x;
      for (int i = r0; i < n_dirty; i += rs) {
        int v = __ldcg(&list[i]);
        int2 d = ldcg_dom(&P.dom[v]);
        if (SMEM) sdom[v] = d;
        bad |= d.x > d.y;
      }
      if (bad) set_failed(c);
      if (SMEM) __syncthreads();
      if (!sweep_now) nprop += expand_dirty_rows<SMEM>(c, cur_buf, n_dirty, cur_epoch);
    }
    if (sweep_now && my_chunks > 0) {
      // ---- the streaming sweep over the static descriptor arrays (ring positions keep
      // counting across sweeps so the mbarrier phases stay consistent)
      if (warp == 0) {
        if (lane == 0) {
          // the ring memory doubles as n-ary staging (generic-proxy writes): order them before
          // the async-proxy writes of the next bulk copies
          if (iter > 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          for (int i = (iter == 0 ? pre_issued : 0); i < my_chunks; ++i) {
            const int q = pipe_pos + i, s = q % kStages;
            if (q >= kStages) mbar_wait(&s_empty[s], ((q / kStages) - 1) & 1);
            producer_issue(P, chunk_of(P, cmap, wid + i * workers), ring + s * kStageBytes, &s_full[s]);
          }
        }
      } else {
        if (iter > 0) aw = load_active_words(P, chunk_of(P, cmap, wid));
        for (int i = 0; i < my_chunks; ++i) {
          const int q = pipe_pos + i, s = q % kStages;
          const Chunk ch = chunk_of(P, cmap, wid + i * workers);
          ActiveWords nxt = aw;
          if (i + 1 < my_chunks) nxt = load_active_words(P, chunk_of(P, cmap, wid + (i + 1) * workers));
          mbar_wait(&s_full[s], (q / kStages) & 1);
          nprop += sweep_consume<SMEM>(c, ch, ring + s * kStageBytes, aw);
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[s]);
          aw = nxt;
        }
      }
      pipe_pos += my_chunks;
    }
    if (iter == 0 && P.trace) { __syncthreads(); trace_mark(P, 3); }
    // n-ary propagators: one CTA each; re-run when one of their operands is dirty.
    if (P.n_nary > 0 && (iter > 0 || P.full_sweep || n_dirty > 0)) {
      for (int s = blockIdx.x; s < P.n_nary; s += gridDim.x) {
        if (!((__ldcg(&P.nary_active[s >> 5]) >> (s & 31)) & 1u)) continue;
        unsigned ev = eval_distinct<SMEM>(c, s, ring, cur_epoch, iter == 0 && P.full_sweep);
        if (threadIdx.x == 0) nprop += ev;
      }
    }

    // block-level propagation count, then the barrier + decision
    for (int o = 16; o; o >>= 1) nprop += __shfl_xor_sync(0xffffffffu, nprop, o);
    if (lane == 0 && nprop) atomicAdd(&s_block_props, nprop);
    nprop = 0;
    __syncthreads();
    unsigned bp = 0;
    if (threadIdx.x == 0) { bp = s_block_props; s_block_props = 0; }
    if (iter == 0) trace_mark(P, 4);
    dec = grid_barrier(P, gen, bp, true, s_flags, iter, next_buf);
    if (iter == 0) trace_mark(P, 5);
    if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && iter < 32) {  // per-iteration record
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.trace[8 * 256 + iter * 4 + 0] = t;
      P.trace[8 * 256 + iter * 4 + 1] = (unsigned long long)*(volatile int*)&ctl->dirty_cnt[next_buf];
      P.trace[8 * 256 + iter * 4 + 2] = dec;
    }
    ++iter;
    if (dec != D_CONTINUE && dec != D_SWEEP) break;
    sweep_now = dec == D_SWEEP;
  }

  trace_mark(P, 6);
  if (blockIdx.x == 0) {
    // Branch::distribute takes a label right after an Unknown node (branch.rs:36-49): leave the
    // copy of the fixpoint domains in the next label slot so that pcp_label is free.
    if (P.snapshot_to && dec == D_FIXPOINT)
      for (int v = threadIdx.x; v < P.V; v += blockDim.x) P.snapshot_to[v] = ldcg_dom(&P.dom[v]);
    if (threadIdx.x == 0) {
      Result r;
      r.failed = dec == D_FAILED;
      r.trail_cnt = *(volatile unsigned*)&ctl->trail_cnt;
      r.iterations = iter;
      r.epoch = epoch0 + iter + 1;
      r.propagations = *(volatile unsigned long long*)&ctl->propagations;
      r.decision = dec;
      *P.result = r;
      ctl->epoch = epoch0 + iter + 1;
      ctl->iterations = iter;
      ctl->last_decision = dec;
      ctl->dirty_cnt[0] = ctl->dirty_cnt[1] = ctl->dirty_cnt[2] = 0;
    }
  }
}

__global__ void pcp_fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// incremental launches: stamp the host-seeded dirty variables with the epoch of iteration 0
__global__ void pcp_seed_dirty_kernel(const int* list, int n, uint32_t* dirty_stamp, const Control* ctl) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dirty_stamp[list[i]] = ctl->epoch;
}

}  // namespace pcpd
