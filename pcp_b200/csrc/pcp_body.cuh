// pcp_body.cuh -- the kernels themselves: per-propagator evaluation, sweeps, n-ary propagators,
// worklist iterations, the per-node fixpoint and the two persistent kernels.  Compiled once per
// kernel variant (pcp_device.cuh): included inside `namespace pcpd { namespace PCP_VARIANT {`,
// so everything in pcp_common.cuh (launch parameters, constants, memory helpers) is visible
// unqualified.  No include guard on purpose.
//
// PCP_BIN_ONLY: the variant for stores whose propagators are all binary (XLessY / XNeqY /
// XEqY: the n-queens and pairwise-distinct stores).  It leaves out the ternary, disjunction
// and n-ary code.  Same source, a third less machine code: the C2 node is sensitive to the
// size of this one big kernel (instruction fetch is its second largest stall class), and a
// same-box A/B of the two variants on C2 measured 19.3 -> 17.4 us per node in the device search.

#include "pcp_eval.cuh"

// The CTA's place among the CTAs that serve its engine.  A launch for one engine: (blockIdx.x,
// gridDim.x).  A batched launch (pcp_fixpoint_batch_kernel) serves one engine per group of
// consecutive CTAs: rank and count are then relative to the group, and so is everything that
// deals work out by CTA or meets at the engine's barrier.  Set by thread 0 at kernel entry,
// read after the first block-wide barrier.
__shared__ unsigned s_cta_rank_, s_cta_count_;
__device__ __forceinline__ unsigned cta_rank() { return s_cta_rank_; }
__device__ __forceinline__ unsigned cta_count() { return s_cta_count_; }
__device__ __forceinline__ void cta_place(unsigned rank, unsigned count) {
  if (threadIdx.x == 0) { s_cta_rank_ = rank; s_cta_count_ = count; }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// iteration 0: the streaming sweep.  A chunk = up to kChunk* propagators of one family;
// chunk g belongs to CTA g % cta_count().  Warp 0 (one lane) is the producer: it arms the
// stage's `full` mbarrier with the byte count and issues the TMA bulk copies; the 31
// consumer warps wait on `full`, evaluate from shared memory and release the stage through
// the `empty` mbarrier.
// ---------------------------------------------------------------------------------------
struct ChunkMap {
  int nch[3];   // chunks per family
  int total;
};
__device__ __forceinline__ constexpr int chunk_props(int fam) { return fam == 0 ? kChunkBin : (fam == 1 ? kChunkTer : kChunkDj); }
__device__ __forceinline__ ChunkMap chunk_map(const Params& P) {
  ChunkMap m;
  const int cb = P.fam[0].cdesc ? kChunkBinC : kChunkBin;
  m.nch[0] = (P.fam[0].n_static + cb - 1) / cb;
  m.nch[1] = (P.fam[1].n_static + kChunkTer - 1) / kChunkTer;
  m.nch[2] = (P.fam[2].n_static + kChunkDj - 1) / kChunkDj;
  m.total = m.nch[0] + m.nch[1] + m.nch[2];
  return m;
}

// mbarrier / bulk-copy helpers on shared-space addresses (the hot loop keeps no generic pointers)
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// One family's share of the sweep for this CTA: chunks g0, g0 + workers, ... (cnt of them),
// ring positions pipe_pos, pipe_pos + 1, ...  Everything the loop needs travels by value: a
// dereference of the kernel parameters from an out-of-line function is a generic load.
struct FamSweep {
  const int4* desc;
  const int2* descB;
  const uint2* cdesc;    // compact binary stream (or nullptr)
  const uint32_t* active;
  const int2* dom;
  int n;                 // propagators covered (n_static)
  int g0, workers, cnt;
  int pipe_pos, first;   // `first` chunks were already issued (pre_issue)
  uint32_t ring_s, full_s, empty_s, sdom_s;
  int have_aw;           // the caller prefetched the active words of the first chunk
};

// Propagators found entailed during a sweep: their active bit is cleared with a fire-and-
// forget reduction (a static propagator is evaluated by exactly one thread per sweep, so
// nobody needs the old bit) and their references are collected in shared memory; the CTA
// appends them to the trail (Store::unlink_prop, propagation/store.rs:200-207) with one
// reservation before the iteration's barrier.  The order inside one node's trail segment is
// immaterial: a restore re-activates the whole suffix (store.rs:319-323).
constexpr int kTrailBuf = 4088;
struct TrailBuf { unsigned n; unsigned pad; unsigned ref[kTrailBuf]; };
__device__ __forceinline__ void sweep_deactivate(const Ctx& c, const FamSweep& a, unsigned fam, int slot) {
  atomicAnd(const_cast<uint32_t*>(&a.active[slot >> 5]), ~(1u << (slot & 31)));
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned base = 0;
  if (lane == leader) asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"(c.tbuf_s), "r"(__popc(m)) : "memory");
  base = __shfl_sync(m, base, leader);
  const unsigned pos = base + __popc(m & lanemask_lt());
  const unsigned ref = make_ref(fam, (unsigned)slot);
  if (pos < (unsigned)kTrailBuf) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.tbuf_s + 8u + 4u * pos), "r"(ref) : "memory");
  } else {  // buffer full: straight to the trail, one reservation per warp
    const unsigned mo = __activemask();
    const int lo = __ffs(mo) - 1;
    unsigned gb = 0;
    if (lane == lo) gb = atomicAdd(c.trail_cnt, (unsigned)__popc(mo));
    gb = __shfl_sync(mo, gb, lo);
    c.trail[gb + __popc(mo & lanemask_lt())] = ref;
  }
}
// Every thread of the CTA, before the iteration's barrier.
__device__ __forceinline__ void trail_flush(const Params& P, TrailBuf* tb) {
  __shared__ unsigned s_base;
  const unsigned n = min(tb->n, (unsigned)kTrailBuf);  // (uniform: read after a barrier)
  if (n == 0) return;
  if (threadIdx.x == 0) s_base = atomicAdd(&P.ctl->trail_cnt, n);
  __syncthreads();
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) P.trail[s_base + i] = tb->ref[i];
  __syncthreads();
  if (threadIdx.x == 0) tb->n = 0u;
}

template <int FAM>
__device__ __forceinline__ void producer_issue(const FamSweep& a, int g, uint32_t stage, uint32_t full) {
  if (FAM == F_BIN && a.cdesc) {  // compact stream: 8 B per descriptor, twice the descriptors per stage
    const int cbase = g * kChunkBinC;
    const unsigned bytes = ((unsigned)min(kChunkBinC, a.n - cbase) * 8u + 15u) & ~15u;
    mbar_expect_tx_s(full, bytes);
    bulk_g2s_s(stage, a.cdesc + cbase, bytes, full);
    return;
  }
  const int base = g * chunk_props(FAM);
  const int cnt = min(chunk_props(FAM), a.n - base);
  if (FAM == F_BIN) {
    unsigned bytes = (unsigned)cnt * 16u;
    mbar_expect_tx_s(full, bytes);
    bulk_g2s_s(stage, a.desc + base, bytes, full);
  } else if (FAM == F_TER) {
    unsigned ba = (unsigned)cnt * 16u, bb = ((unsigned)cnt * 8u + 15u) & ~15u;
    mbar_expect_tx_s(full, ba + bb);
    bulk_g2s_s(stage, a.desc + base, ba, full);
    bulk_g2s_s(stage + kTerPlaneB, a.descB + base, bb, full);
  } else {
    unsigned bytes = (unsigned)cnt * 48u;
    mbar_expect_tx_s(full, bytes);
    bulk_g2s_s(stage, a.desc + 3 * (size_t)base, bytes, full);
  }
}

__device__ __forceinline__ int4 lds128(uint32_t addr) {
  int4 r;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// Cold path of the sweep: the propagator prunes, fails, is entailed, or has a Constant / Sum
// operand.  The descriptor is re-read from the ring stage (still owned by this warp), so the
// call carries three registers and the hot loop keeps nothing alive for it.
template <bool SMEM>
__device__ __noinline__ void sweep_slow_bin(const Ctx& c, int slot, uint32_t desc_s) {
  const int4 d = lds128(desc_s);
  const IV x = rd<SMEM>(c, dec_var28((unsigned)d.x), d.y), y = rd<SMEM>(c, d.z, d.w);
#ifdef PCP_SET
  if (!bin_is_noop_set(c, d, x, y)) eval_full_bin(c, slot, d, x, y);
#else
  if (!bin_is_noop((unsigned)d.x >> 28, x, y)) eval_full_bin(c, slot, d, x, y);
#endif
}
template <bool SMEM>
__device__ __noinline__ void sweep_slow_ter(const Ctx& c, int slot, uint32_t a_s, uint32_t b_s) {
  const int4 a = lds128(a_s);
  const int2 b = lds_dom(b_s);
  const IV x = rd<SMEM>(c, dec_var28((unsigned)a.x), a.y), y = rd<SMEM>(c, a.z, a.w), z = rd<SMEM>(c, b.x, b.y);
  if (!ter_is_noop((unsigned)a.x >> 28, x, y, z)) eval_full_ter(c, slot, a, b, x, y, z);
}
template <bool SMEM>
__device__ __forceinline__ int2 rd_plain(uint32_t sdom_s, const int2* dom, int var) {
  return SMEM ? lds_dom(sdom_s + 8u * (unsigned)var) : ldcg_dom(&dom[var]);
}

// The words of the `active` bit set a consumer warp needs for one chunk: its descriptors are
// consecutive and start at a multiple of their count, so the words are one aligned vector
// load; loaded one chunk ahead so that the L2 round trip is off the critical path of the sweep.
struct ActiveWords { unsigned w[4]; };
struct ActiveWords8 { unsigned w[8]; };
__device__ __forceinline__ ActiveWords8 load_active8(const uint32_t* active, int base, int n) {
  ActiveWords8 a;
#pragma unroll
  for (int i = 0; i < 8; ++i) a.w[i] = 0u;
  if (base < n) {
    const uint4* p = reinterpret_cast<const uint4*>(active + (base >> 5));
    const uint4 v0 = __ldcg(p), v1 = __ldcg(p + 1);
    a.w[0] = v0.x; a.w[1] = v0.y; a.w[2] = v0.z; a.w[3] = v0.w;
    a.w[4] = v1.x; a.w[5] = v1.y; a.w[6] = v1.z; a.w[7] = v1.w;
  }
  return a;
}
template <int G>
__device__ __forceinline__ ActiveWords load_active(const uint32_t* active, int base, int n) {
  ActiveWords a;
  a.w[0] = a.w[1] = a.w[2] = a.w[3] = 0u;
  if (base < n) {
    if (G == 4) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(active + (base >> 5)));
      a.w[0] = v.x; a.w[1] = v.y; a.w[2] = v.z; a.w[3] = v.w;
    } else {
      const uint2 v = __ldcg(reinterpret_cast<const uint2*>(active + (base >> 5)));
      a.w[0] = v.x; a.w[1] = v.y;
    }
  }
  return a;
}
// bits [0, rem) of a 32-propagator group that starts `rem` descriptors before the end
__device__ __forceinline__ unsigned tail_mask(int rem) {
  return rem >= 32 ? 0xffffffffu : (rem > 0 ? (1u << rem) - 1u : 0u);
}

// Producer side of a family sweep: one lane keeps the TMA ring full.
template <int FAM>
__device__ __forceinline__ void sweep_produce(const FamSweep& a) {
  for (int i = a.first; i < a.cnt; ++i) {
    const int q = a.pipe_pos + i, s = q % kStages;
    if (q >= kStages) mbar_wait_s(a.empty_s + 8u * s, ((q / kStages) - 1) & 1);
    producer_issue<FAM>(a, a.g0 + i * a.workers, a.ring_s + (unsigned)(s * kStageBytes), a.full_s + 8u * s);
  }
}

// Ring position of a consumer warp, carried across the chunks of a family sweep (no division
// or re-derivation per chunk): stage index, phase parity of its `full` barrier, and the
// shared-space addresses that go with them.
struct RingPos {
  unsigned s, ph;
  uint32_t stage, full, empty;
};
__device__ __forceinline__ RingPos ring_pos(const FamSweep& a) {
  RingPos r;
  r.s = (unsigned)a.pipe_pos % kStages;
  r.ph = ((unsigned)a.pipe_pos / kStages) & 1u;
  r.stage = a.ring_s + r.s * (unsigned)kStageBytes;
  r.full = a.full_s + 8u * r.s;
  r.empty = a.empty_s + 8u * r.s;
  return r;
}
__device__ __forceinline__ void ring_advance(RingPos& r) {
  ++r.s; r.stage += (unsigned)kStageBytes; r.full += 8u; r.empty += 8u;
  if (r.s == (unsigned)kStages) {
    r.s = 0u; r.ph ^= 1u; r.stage -= (unsigned)kRingBytes; r.full -= 8u * kStages; r.empty -= 8u * kStages;
  }
}

// XEqYPlusZ over plain variables that prunes (the common case of the first sweeps of an
// arithmetic store): propagate + is_subsumed inline, exactly eval_ter's T_EQ branch followed
// by finish_eval -- geq half, then leq half on the narrowed values (x_eq_y_plus_z.rs:79-81),
// Kleene entailment on the result, nothing applied on failure -- with the updates issued as
// fire-and-forget reductions straight from registers.  Entailment (rare) stays out of line.
__device__ __forceinline__ bool sweep_upd(const Ctx& c, int var, int off, IV o, IV n) {
  const bool lo = n.lo > o.lo, hi = n.hi < o.hi;
#ifndef PCP_EXP_NO_RED
  if (lo) atomicMax(&c.dom_w[var].x, n.lo - off);
  if (hi) atomicMin(&c.dom_w[var].y, n.hi - off);
#endif
  // the variable is queued: in the CTA's bitmap (one coalesced flush into the dirty set after
  // the sweep instead of a scattered reduction per narrowing), else directly.  (Looking the bit
  // up first was measured slower: the look is a dependent load on the update path.)
#ifndef PCP_EXP_NO_MARK
  if (lo || hi) {
    const uint32_t dbm_s = c.dbm_s;
    if (dbm_s) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(dbm_s + 4u * (unsigned)(var >> 5)), "r"(1u << (var & 31)) : "memory");
    else atomicOr(&c.next_bits[var >> 5], 1u << (var & 31));
  }
#endif
  return lo || hi;
}
__device__ __forceinline__ void sweep_ter_eq_update(const Ctx& c, const FamSweep& a, int slot, int4 d, int2 e, IV x0, IV y0, IV z0) {
#ifdef PCP_SET
  eval_full_ter(c, slot, d, e, x0, y0, z0);  // (bounds land on set elements: the staged path knows how)
  return;
#endif
  IV x = x0, y = y0, z = z0;
  const bool ok = prop_greater(x, y, z, 0) && prop_less(x, y, z, 0);
  const int s = ok ? sub_eq(x, y, z) : -1;
  if (s < 0) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.flags_s + 4u), "r"(1) : "memory");  // flags[1]: failure
    return;
  }
  bool ch = sweep_upd(c, (int)((unsigned)d.x & kConstVar28), d.y, x0, x);
  ch |= sweep_upd(c, d.z, d.w, y0, y);
  ch |= sweep_upd(c, e.x, e.y, z0, z);
  if (ch) asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.flags_s), "r"(1) : "memory");  // flags[0]: narrowed
  if (s > 0) sweep_deactivate(c, a, F_TER, slot);
}

// XNeqY over plain variables that is not a no-op: eval_bin's B_NEQ branch + finish_eval inline
// (x_neq_y.rs:82-93 on Interval: a singleton side trims the matching bound of the other side;
// is_subsumed = not XEqY, x_neq_y.rs:71-73).
#ifdef PCP_SET
// IntervalSet: the assigned side's value leaves the other domain wherever it sits (a bound moves to
// the next value still in the set, an interior value is cleared), after which the two are
// disjoint and the propagator is entailed; two unassigned sides are entailed when their sets are
// disjoint (decided exactly here: the hot loop only found no witness at the larger lower bound).
__device__ __noinline__ void sweep_neq_update(const Ctx& c, const FamSweep& a, int slot, int4 d, IV x, IV y) {
  const SetW sw = set_of(*c.P);
  const int xv = (int)((unsigned)d.x & kConstVar28), yv = d.z;
  const bool xs = x.lo == x.hi, ys = y.lo == y.hi;
  if (xs || ys) {
    const int v = xs ? x.lo : y.lo;                    // the assigned side's value (view space)
    const int tv = xs ? yv : xv, to = xs ? d.w : d.y;  // the other side
    const IV t = xs ? y : x;
    if (v >= t.lo && v <= t.hi) {
      if (t.lo == t.hi) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.flags_s + 4u), "r"(1) : "memory"); return; }
      bool ch;
      if (v == t.lo) {
        const int r = sb_next(sw, tv, v - to + 1, t.hi - to);
        if (r == INT32_MAX) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.flags_s + 4u), "r"(1) : "memory"); return; }
        ch = sweep_upd(c, tv, to, t, IV{r + to, t.hi});
      } else if (v == t.hi) {
        const int r = sb_prev(sw, tv, v - to - 1, t.lo - to);
        if (r == INT32_MIN) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.flags_s + 4u), "r"(1) : "memory"); return; }
        ch = sweep_upd(c, tv, to, t, IV{t.lo, r + to});
      } else {
        ch = sb_clear(sw, tv, v - to);
        if (ch) {
          const uint32_t dbm_s = c.dbm_s;
          if (dbm_s) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(dbm_s + 4u * (unsigned)(tv >> 5)), "r"(1u << (tv & 31)) : "memory");
          else atomicOr(&c.next_bits[tv >> 5], 1u << (tv & 31));
        }
      }
      if (ch) asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.flags_s), "r"(1) : "memory");
    }
    sweep_deactivate(c, a, F_BIN, slot);
    return;
  }
  if (x.hi < y.lo || y.hi < x.lo) { sweep_deactivate(c, a, F_BIN, slot); return; }
  if (sb_witness(sw, xv, x.lo - d.y, x.hi - d.y, d.y, yv, y.lo - d.w, y.hi - d.w, d.w)) return;
  if (sb_disjoint(sw, xv, x.lo - d.y, x.hi - d.y, d.y, yv, y.lo - d.w, y.hi - d.w, d.w)) sweep_deactivate(c, a, F_BIN, slot);
}
#else
__device__ __forceinline__ void sweep_neq_update(const Ctx& c, const FamSweep& a, int slot, int4 d, IV x, IV y) {
  IV nx = x, ny = y;
  if (x.lo == x.hi) {
    if (ny.lo == x.lo) ny.lo++; else if (ny.hi == x.lo) ny.hi--;
  } else if (y.lo == y.hi) {
    if (nx.lo == y.lo) nx.lo++; else if (nx.hi == y.lo) nx.hi--;
  }
  const bool fail = ny.lo > ny.hi || nx.lo > nx.hi || (nx.lo == ny.hi && nx.hi == ny.lo);
  if (fail) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.flags_s + 4u), "r"(1) : "memory");
    return;
  }
  bool ch = sweep_upd(c, d.z, d.w, y, ny);
  ch |= sweep_upd(c, (int)((unsigned)d.x & kConstVar28), d.y, x, nx);
  if (ch) asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.flags_s), "r"(1) : "memory");
  if (nx.hi < ny.lo || ny.hi < nx.lo) sweep_deactivate(c, a, F_BIN, slot);
}
#endif

// The streaming sweep of one CTA over the binary family (XLessY / XNeqY / XEqY): warp 0 (one
// lane) keeps the ring full, every other warp owns 128 consecutive descriptors of each chunk
// (kGroupsBin groups of 32 = four words of the `active` set).  Hot loop: the descriptor loads
// of all groups first, then the eight domain reads, then the tests; whatever is not a no-op
// goes out of line with three registers.  NEQ_PLAIN: every static descriptor of the family
// is an XNeqY over plain variables (host-side flags Family::all_plain / kind_mask) -- the
// n-queens and pairwise-distinct stores -- so neither the kind nor the operand encoding is
// inspected.  Returns the number of evaluations of the warp (the same value in every lane).
template <bool SMEM, bool NEQ_PLAIN>
__device__ __noinline__ unsigned sweep_bin(const Ctx& c, const FamSweep a, uint4 aw4) {
  ActiveWords aw;
  aw.w[0] = aw4.x; aw.w[1] = aw4.y; aw.w[2] = aw4.z; aw.w[3] = aw4.w;
  constexpr int G = kGroupsBin;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (lane == 0) sweep_produce<F_BIN>(a);
    return 0u;
  }
  const unsigned lane_bit = 1u << lane;
  const uint32_t lane_off = (uint32_t)((warp - 1) * (32 * G) + lane) * 16u;
#ifdef PCP_SET
  const SetW sw = set_of(*c.P);
#endif
  const int stride = a.workers * kChunkBin;
  int base = a.g0 * kChunkBin + (warp - 1) * (32 * G);  // this warp's first descriptor of the chunk
  if (!a.have_aw) aw = load_active<G>(a.active, base, a.n);
  RingPos r = ring_pos(a);
  unsigned nprop = 0;
  for (int i = 0; i < a.cnt; ++i) {
    const int nbase = base + stride;
    ActiveWords nxt;
    nxt.w[0] = nxt.w[1] = nxt.w[2] = nxt.w[3] = 0u;
    if (i + 1 < a.cnt) nxt = load_active<G>(a.active, nbase, a.n);
    const int rem = a.n - base;
    if (rem < 32 * G) {  // last chunk of the family
#pragma unroll
      for (int g = 0; g < G; ++g) aw.w[g] &= tail_mask(rem - 32 * g);
    }
    const uint32_t d_s = r.stage + lane_off;
    mbar_wait_s(r.full, r.ph);
    int4 d[G];
    int2 dx[G], dy[G];
    bool on[G], plain[G], need[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      nprop += __popc(aw.w[g]);
      on[g] = (aw.w[g] & lane_bit) != 0u;
      d[g] = make_int4(0, 0, 0, 0);
      if (on[g]) d[g] = lds128(d_s + 512u * g);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const unsigned xv = (unsigned)d[g].x & kConstVar28;
      plain[g] = NEQ_PLAIN || (xv < kSumBase28 && d[g].z >= 0);
      dx[g] = rd_plain<SMEM>(a.sdom_s, a.dom, plain[g] ? (int)xv : 0);
      dy[g] = rd_plain<SMEM>(a.sdom_s, a.dom, plain[g] ? d[g].z : 0);
    }
    bool any = false;
#ifdef PCP_SET
    // IntervalSet: two unassigned overlapping views make an XNeqY a no-op only if they share a value.
    // One bit test proves it in the common case (dense domains): the larger lower bound belongs to
    // its own set, is it in the other one?  The probes of all groups are in flight together.
    unsigned pw[G], pbit[G];
    bool cand[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const IV x{dx[g].x + d[g].y, dx[g].y + d[g].y}, y{dy[g].x + d[g].w, dy[g].y + d[g].w};
      const unsigned kind = NEQ_PLAIN ? (unsigned)B_NEQ : (unsigned)d[g].x >> 28;
      cand[g] = on[g] && plain[g] && kind == B_NEQ && x.lo != x.hi && y.lo != y.hi && !(x.hi < y.lo || y.hi < x.lo);
      pw[g] = 0u;
      pbit[g] = 0u;
      if (cand[g]) {
        const bool lx = x.lo >= y.lo;
        const int pv = lx ? d[g].z : (int)((unsigned)d[g].x & kConstVar28);
        const unsigned b = (unsigned)((lx ? x.lo - d[g].w : y.lo - d[g].y) - sw.base);
        pw[g] = __ldcg(sw.bits + (size_t)pv * (unsigned)sw.W + (b >> 5));
        pbit[g] = b & 31u;
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const IV x{dx[g].x + d[g].y, dx[g].y + d[g].y}, y{dy[g].x + d[g].w, dy[g].y + d[g].w};
      const unsigned kind = NEQ_PLAIN ? (unsigned)B_NEQ : (unsigned)d[g].x >> 28;
      const bool noop = cand[g] ? ((pw[g] >> pbit[g]) & 1u) != 0u : (plain[g] && kind == B_LESS && bin_is_noop(B_LESS, x, y));
      need[g] = on[g] && !noop;
      any |= need[g];
    }
#else
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const IV x{dx[g].x + d[g].y, dx[g].y + d[g].y}, y{dy[g].x + d[g].w, dy[g].y + d[g].w};
      need[g] = on[g] && !(plain[g] && bin_is_noop(NEQ_PLAIN ? (unsigned)B_NEQ : (unsigned)d[g].x >> 28, x, y));
      any |= need[g];
    }
#endif
    if (any) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (!need[g]) continue;
        if (NEQ_PLAIN) {
          const IV x{dx[g].x + d[g].y, dx[g].y + d[g].y}, y{dy[g].x + d[g].w, dy[g].y + d[g].w};
          sweep_neq_update(c, a, base + 32 * g + lane, d[g], x, y);
        } else {
          sweep_slow_bin<SMEM>(c, base + 32 * g + lane, d_s + 512u * g);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_s(r.empty);
    ring_advance(r);
    aw = nxt;
    base = nbase;
  }
  return nprop;
}

// The lean binary sweep over the compact stream (Family::cdesc): every static descriptor is an
// XNeqY over plain variables and fits 8 bytes, so a 32 KB stage holds 3840 of them and every
// consumer warp owns 256 consecutive ones (eight groups = eight words of the `active` set) per
// chunk -- half the bytes from HBM / L2 per propagation and half the per-chunk overhead.
template <bool SMEM>
__device__ __noinline__ unsigned sweep_bin_compact(const Ctx& c, const FamSweep a, uint4 aw4a, uint4 aw4b) {
  constexpr int G = kGroupsBinC;
  ActiveWords8 aw;
  aw.w[0] = aw4a.x; aw.w[1] = aw4a.y; aw.w[2] = aw4a.z; aw.w[3] = aw4a.w;
  aw.w[4] = aw4b.x; aw.w[5] = aw4b.y; aw.w[6] = aw4b.z; aw.w[7] = aw4b.w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (lane == 0) sweep_produce<F_BIN>(a);
    return 0u;
  }
  const unsigned lane_bit = 1u << lane;
  const uint32_t lane_off = (uint32_t)((warp - 1) * (32 * G) + lane) * 8u;
  const int stride = a.workers * kChunkBinC;
  int base = a.g0 * kChunkBinC + (warp - 1) * (32 * G);  // this warp's first descriptor of the chunk
  if (!a.have_aw) aw = load_active8(a.active, base, a.n);
  RingPos r = ring_pos(a);
  unsigned nprop = 0;
  for (int i = 0; i < a.cnt; ++i) {
    const int nbase = base + stride;
    ActiveWords8 nxt;
#pragma unroll
    for (int g = 0; g < G; ++g) nxt.w[g] = 0u;
    if (i + 1 < a.cnt) nxt = load_active8(a.active, nbase, a.n);
    const int rem = a.n - base;
    if (rem < 32 * G) {  // last chunk of the family
#pragma unroll
      for (int g = 0; g < G; ++g) aw.w[g] &= tail_mask(rem - 32 * g);
    }
    const uint32_t d_s = r.stage + lane_off;
    mbar_wait_s(r.full, r.ph);
#pragma unroll
    for (int g = 0; g < G; ++g) nprop += __popc(aw.w[g]);
    // two half-batches of four groups: the register budget of the 16-byte loop, half its
    // per-chunk overhead
#pragma unroll
    for (int h = 0; h < G; h += 4) {
      int2 d[4], dx[4], dy[4];
      unsigned need = 0u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        d[g] = make_int2(0, 0);
        if (aw.w[h + g] & lane_bit) d[g] = lds_dom(d_s + 256u * (h + g));
      }
      // word 0: the byte offsets of the two domains (variable index * 8); word 1: x's view offset and
      // the DIFFERENCE of the two view offsets.  The no-op test of an XNeqY -- both sides unassigned,
      // views overlapping (bin_is_noop) -- needs no view arithmetic then: x.hi < y.lo  <=>
      // dx.hi - dy.lo < yoff - xoff, y.hi < x.lo  <=>  dx.lo - dy.hi > yoff - xoff.
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const unsigned bx = (unsigned)d[g].x & 0xffffu, by = (unsigned)d[g].x >> 16;
        dx[g] = SMEM ? lds_dom(a.sdom_s + bx) : ldcg_dom(reinterpret_cast<const int2*>(reinterpret_cast<const char*>(a.dom) + bx));
        dy[g] = SMEM ? lds_dom(a.sdom_s + by) : ldcg_dom(reinterpret_cast<const int2*>(reinterpret_cast<const char*>(a.dom) + by));
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int delta = d[g].y >> 16;
        const bool noop = dx[g].x != dx[g].y && dy[g].x != dy[g].y && !(dx[g].y - dy[g].x < delta || dx[g].x - dy[g].y > delta);
        if ((aw.w[h + g] & lane_bit) && !noop) need |= 1u << g;
      }
      if (need) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (!((need >> g) & 1u)) continue;
          const int xo = (int)(short)((unsigned)d[g].y & 0xffffu), yo = xo + (d[g].y >> 16);
          const IV x{dx[g].x + xo, dx[g].y + xo}, y{dy[g].x + yo, dy[g].y + yo};
          const int4 full = make_int4((int)((B_NEQ << 28) | (((unsigned)d[g].x & 0xffffu) >> 3)), xo, (int)((unsigned)d[g].x >> 19), yo);
          sweep_neq_update(c, a, base + 32 * (h + g) + lane, full, x, y);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_s(r.empty);
    ring_advance(r);
#pragma unroll
    for (int g = 0; g < G; ++g) aw.w[g] = nxt.w[g];
    base = nbase;
  }
  return nprop;
}

// The same for the ternary family (XGreaterYPlusZ / XLessYPlusZ / XEqYPlusZ): kGroupsTer
// groups per warp and chunk, the (x, y) plane and the z plane of the stage read separately.
// EQ_PLAIN: every static descriptor is an XEqYPlusZ over plain variables.
template <bool SMEM, bool EQ_PLAIN>
__device__ __noinline__ unsigned sweep_ter(const Ctx& c, const FamSweep a, uint4 aw4) {
  ActiveWords aw;
  aw.w[0] = aw4.x; aw.w[1] = aw4.y; aw.w[2] = aw4.z; aw.w[3] = aw4.w;
  constexpr int G = kGroupsTer;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (lane == 0) sweep_produce<F_TER>(a);
    return 0u;
  }
  const unsigned lane_bit = 1u << lane;
  const uint32_t lane_j = (uint32_t)((warp - 1) * (32 * G) + lane);
  const int stride = a.workers * kChunkTer;
  int base = a.g0 * kChunkTer + (warp - 1) * (32 * G);
  if (!a.have_aw) aw = load_active<G>(a.active, base, a.n);
  RingPos r = ring_pos(a);
  unsigned nprop = 0;
  for (int i = 0; i < a.cnt; ++i) {
    const int nbase = base + stride;
    ActiveWords nxt;
    nxt.w[0] = nxt.w[1] = nxt.w[2] = nxt.w[3] = 0u;
    if (i + 1 < a.cnt) nxt = load_active<G>(a.active, nbase, a.n);
    const int rem = a.n - base;
    if (rem < 32 * G) {
#pragma unroll
      for (int g = 0; g < G; ++g) aw.w[g] &= tail_mask(rem - 32 * g);
    }
    const uint32_t a_s = r.stage + 16u * lane_j, b_s = r.stage + (unsigned)kTerPlaneB + 8u * lane_j;
    mbar_wait_s(r.full, r.ph);
    int4 d[G];
    int2 e[G], dx[G], dy[G], dz[G];
    bool on[G], plain[G], need[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      nprop += __popc(aw.w[g]);
      on[g] = (aw.w[g] & lane_bit) != 0u;
      d[g] = make_int4(0, 0, 0, 0);
      e[g] = make_int2(0, 0);
      if (on[g]) { d[g] = lds128(a_s + 512u * g); e[g] = lds_dom(b_s + 256u * g); }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const unsigned xv = (unsigned)d[g].x & kConstVar28;
      plain[g] = EQ_PLAIN || (xv < kSumBase28 && d[g].z >= 0 && e[g].x >= 0);
      dx[g] = rd_plain<SMEM>(a.sdom_s, a.dom, plain[g] ? (int)xv : 0);
      dy[g] = rd_plain<SMEM>(a.sdom_s, a.dom, plain[g] ? d[g].z : 0);
      dz[g] = rd_plain<SMEM>(a.sdom_s, a.dom, plain[g] ? e[g].x : 0);
    }
    bool any = false;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const IV x{dx[g].x + d[g].y, dx[g].y + d[g].y}, y{dy[g].x + d[g].w, dy[g].y + d[g].w},
               z{dz[g].x + e[g].y, dz[g].y + e[g].y};
      need[g] = on[g] && !(plain[g] && ter_is_noop(EQ_PLAIN ? (unsigned)T_EQ : (unsigned)d[g].x >> 28, x, y, z));
      any |= need[g];
    }
    if (any) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (!need[g]) continue;
        if (EQ_PLAIN) {
          const IV x{dx[g].x + d[g].y, dx[g].y + d[g].y}, y{dy[g].x + d[g].w, dy[g].y + d[g].w},
                   z{dz[g].x + e[g].y, dz[g].y + e[g].y};
          sweep_ter_eq_update(c, a, base + 32 * g + lane, d[g], e[g], x, y, z);
        } else {
          sweep_slow_ter<SMEM>(c, base + 32 * g + lane, a_s + 512u * g, b_s + 256u * g);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_s(r.empty);
    ring_advance(r);
    aw = nxt;
    base = nbase;
  }
  return nprop;
}

// The disjunction family (logic/disjunction.rs) is small wherever it occurs: one generic loop.
template <bool SMEM>
__device__ __noinline__ unsigned sweep_dj(const Ctx& c, const FamSweep a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (lane == 0) sweep_produce<F_DJ>(a);
    return 0u;
  }
  const int cw = warp - 1;
  unsigned nprop = 0;
  int base = a.g0 * kChunkDj;
  for (int i = 0; i < a.cnt; ++i) {
    const int q = a.pipe_pos + i, s = q % kStages;
    const int cnt = min(kChunkDj, a.n - base);
    const uint32_t stage = a.ring_s + (unsigned)(s * kStageBytes);
    mbar_wait_s(a.full_s + 8u * s, (q / kStages) & 1);
    for (int j0 = cw * 32; j0 < cnt; j0 += kConsumerWarps * 32) {
      const int j = j0 + lane;
      const unsigned word = __ldcg(&a.active[(base + j0) >> 5]);  // base, j0 are multiples of 32
      const bool on = j < cnt && ((word >> lane) & 1u);
      nprop += __popc(__ballot_sync(0xffffffffu, on));
      if (on) {
        const int4 q0 = lds128(stage + 48u * (unsigned)j), q1 = lds128(stage + 48u * (unsigned)j + 16u),
                   q2 = lds128(stage + 48u * (unsigned)j + 32u);
        if (!dj_is_noop<SMEM>(c, q0, q1, q2)) eval_full_dj<SMEM>(c, base + j, q0, q1, q2);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_s(a.empty_s + 8u * s);
    base += a.workers * kChunkDj;
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
__device__ __forceinline__ unsigned eval_distinct(const Params& P, const Ctx& c, int slot, char* smem_nary, const uint32_t* cur_bits,
                                                  bool unconditional) {
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
      if (!unconditional && ((__ldcg(&cur_bits[op.x >> 5]) >> (op.x & 31)) & 1u)) any_dirty = 1;
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

#ifdef PCP_SET
// ---------------------------------------------------------------------------------------
// n-ary Distinct on IntervalSet domains: the same conjunction of pairwise XNeqY
// (distinct.rs:69-126), whose fixpoint on sets is: every assigned operand's value leaves the
// set of every other operand -- bounds and interior alike -- an operand that becomes assigned
// joins in, two equal assigned operands fail.  Entailed iff the operands' sets are pairwise
// disjoint (Conjunction::is_subsumed, conjunction.rs:77-94, over XNeqY::is_subsumed).
// smem (the idle ring): int2 ops[k]; int2 iv[k]; int tab[tabsz]; int newv[k].
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool views_disjoint(const SetW& sw, int2 oa, int2 a, int2 ob, int2 b) {
  if (a.y < b.x || b.y < a.x) return true;
  if (a.x == a.y || oa.x < 0) {           // a is a single value
    if (b.x == b.y || ob.x < 0) return false;   // (the bounds overlap: the same value)
    return !sb_test(sw, ob.x, a.x - ob.y);
  }
  if (b.x == b.y || ob.x < 0) return !sb_test(sw, oa.x, b.x - oa.y);
  return sb_disjoint(sw, oa.x, a.x - oa.y, a.y - oa.y, oa.y, ob.x, b.x - ob.y, b.y - ob.y, ob.y);
}
template <bool SMEM>
__device__ __forceinline__ unsigned eval_distinct_set(const Params& P, const Ctx& c, int slot, char* smem_nary, const uint32_t* cur_bits,
                                                      bool unconditional) {
  __shared__ int s_nnew;
  const SetW sw = set_of(P);
  const int b = __ldg(&P.nary_ptr[slot]), e = __ldg(&P.nary_ptr[slot + 1]);
  const int k = e - b;
  int2* ops = reinterpret_cast<int2*>(smem_nary);
  int2* iv = ops + P.nary_max_k;
  int* tab = reinterpret_cast<int*>(iv + P.nary_max_k);
  unsigned tabsz = 4;
  while (tabsz < 2u * (unsigned)k) tabsz <<= 1;
  const unsigned mask = tabsz - 1;
  int* newv = tab + tabsz;
  __syncthreads();
  int any_dirty = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    int2 op = __ldg(&P.nary_ops[b + i]);
    ops[i] = op;
    if (op.x >= 0) {
      int2 d = SMEM ? c.sdom[op.x] : ldcg_dom(&P.dom[op.x]);
      iv[i] = make_int2(d.x + op.y, d.y + op.y);
      if (!unconditional && ((__ldcg(&cur_bits[op.x >> 5]) >> (op.x & 31)) & 1u)) any_dirty = 1;
    } else {
      iv[i] = make_int2(op.y, op.y);
    }
  }
  for (unsigned i = threadIdx.x; i < tabsz; i += blockDim.x) tab[i] = kHashEmpty;
  if (!unconditional) {
    if (!__syncthreads_or(any_dirty)) return 0;
  } else {
    __syncthreads();
  }
  unsigned inserted_bits = 0, inner_bits = 0;  // per owned operand j (k <= 32 * blockDim)
  while (true) {
    if (threadIdx.x == 0) s_nnew = 0;
    __syncthreads();
    int fail = 0, j = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x, ++j) {
      const int2 d = iv[i];
      if (d.x == d.y && !((inserted_bits >> j) & 1u)) {
        inserted_bits |= 1u << j;
        if (!hs_insert(tab, mask, d.x)) fail = 1;            // two equal assigned operands
        else newv[atomicAdd(&s_nnew, 1)] = d.x;
      }
    }
    if (__syncthreads_or(fail)) { set_failed(c); return 1; }
    const int n = s_nnew;
    if (n == 0) break;
    j = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x, ++j) {
      const int2 op = ops[i];
      int2 d = iv[i];
      if (op.x < 0 || d.x == d.y) continue;
      bool moved = false;
      for (int q = 0; q < n && d.x <= d.y; ++q) {
        const int v = newv[q];
        if (v < d.x || v > d.y) continue;
        if (v == d.x) {
          const int r = sb_next(sw, op.x, v - op.y + 1, d.y - op.y);
          d.x = r == INT32_MAX ? d.y + 1 : r + op.y;
          moved = true;
        } else if (v == d.y) {
          const int r = sb_prev(sw, op.x, v - op.y - 1, d.x - op.y);
          d.y = r == INT32_MIN ? d.x - 1 : r + op.y;
          moved = true;
        } else if (sb_clear(sw, op.x, v - op.y)) {
          inner_bits |= 1u << j;
        }
      }
      if (d.x > d.y) { fail = 1; continue; }
      if (moved) iv[i] = d;
    }
    if (__syncthreads_or(fail)) { set_failed(c); return 1; }
  }
  // write back: bounds through the staged path, interior removals are already in the sets
  int j = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x, ++j) {
    const int2 op = ops[i];
    if (op.x < 0) continue;
    const int2 d = iv[i];
    const int2 cur = ldcg_dom(&P.dom[op.x]);
    const IV curv{cur.x + op.y, cur.y + op.y};
    if (!tighten(c, op.x, op.y, curv, max(curv.lo, d.x), min(curv.hi, d.y))) set_failed(c);
    if (((inner_bits >> j) & 1u) && c.mark_dirty) {
      atomicOr(&c.next_bits[op.x >> 5], 1u << (op.x & 31));
      c.flags[0] = 1;
    }
  }
  __syncthreads();
  bool entailed = k <= 1;
  if (k > 1) {
    int overlap = 0;
    for (int i = threadIdx.x; i < k && !overlap; i += blockDim.x) {
      const int2 oa = ops[i], a = iv[i];
      for (int q = i + 1; q < k; ++q)
        if (!views_disjoint(sw, oa, a, ops[q], iv[q])) { overlap = 1; break; }
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
#endif  // PCP_SET

// ---------------------------------------------------------------------------------------
// n-ary AllEqual (propagators/all_equal.rs:47-103 = Conjunction of XEqY(v_i, v_{i+1}),
// cmp/x_eq_y.rs:84-107).  Every XEqY replaces both sides by their intersection, so the chain's
// fixpoint gives every operand the intersection I of all (view-space) domains -- a Constant
// operand takes part as the singleton it is -- and fails iff I is empty; the conjunction is
// entailed iff every pair is an equal pair of singletons, i.e. I is a singleton (an array of
// one variable is the empty conjunction: entailed).  One CTA: block-wide max of the lower and
// min of the upper bounds, then the write-back.  Returns 1 if evaluated (thread 0 only).
// ---------------------------------------------------------------------------------------
template <bool SMEM>
__device__ __forceinline__ unsigned eval_all_equal(const Params& P, const Ctx& c, int slot, const uint32_t* cur_bits,
                                                   bool unconditional) {
  __shared__ int s_lo, s_hi;
  const int b = __ldg(&P.nary_ptr[slot]), e = __ldg(&P.nary_ptr[slot + 1]);
  const int k = e - b;
  __syncthreads();
  if (threadIdx.x == 0) { s_lo = INT32_MIN; s_hi = INT32_MAX; }
  int any_dirty = 0, lo = INT32_MIN, hi = INT32_MAX;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const int2 op = __ldg(&P.nary_ops[b + i]);
    int2 d = make_int2(op.y, op.y);
    if (op.x >= 0) {
      const int2 raw = SMEM ? c.sdom[op.x] : ldcg_dom(&P.dom[op.x]);
      d = make_int2(raw.x + op.y, raw.y + op.y);
      if (!unconditional && ((__ldcg(&cur_bits[op.x >> 5]) >> (op.x & 31)) & 1u)) any_dirty = 1;
    }
    lo = max(lo, d.x);
    hi = min(hi, d.y);
  }
  if (!unconditional) {
    if (!__syncthreads_or(any_dirty)) return 0;  // nothing it depends on changed (all_equal.rs:96-103)
  } else {
    __syncthreads();
  }
  for (int o = 16; o; o >>= 1) {
    lo = max(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = min(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(&s_lo, lo); atomicMin(&s_hi, hi); }
  __syncthreads();
  lo = s_lo;
  hi = s_hi;
  if (k > 1 && lo > hi) {  // two operands are disjoint: some XEqY of the chain fails
    if (threadIdx.x == 0) set_failed(c);
    return 1;
  }
  if (k > 1) {
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const int2 op = __ldg(&P.nary_ops[b + i]);
      if (op.x < 0) continue;  // a Constant inside I needs no update (term/constant.rs:43-53)
      const int2 raw = ldcg_dom(&P.dom[op.x]);
      const IV cur{raw.x + op.y, raw.y + op.y};
      if (!tighten(c, op.x, op.y, cur, max(cur.lo, lo), min(cur.hi, hi))) set_failed(c);
    }
  }
  if (threadIdx.x == 0 && (k <= 1 || lo == hi)) {
    const unsigned bit = 1u << (slot & 31);
    const unsigned old = atomicAnd(&P.nary_active[slot >> 5], ~bit);
    if (old & bit) P.trail[atomicAdd(&P.ctl->trail_cnt, 1u)] = make_ref(F_NARY, (unsigned)slot);
  }
  return 1;
}

#if !defined(PCP_BIN_ONLY) && !defined(PCP_SET)
// ---------------------------------------------------------------------------------------
// Formula trees: one propagator = a tree of Conjunction / Disjunction / Boolean / BooleanNeg over
// cmp leaves, evaluated by one thread on a private copy of the domains of the tree's variables
// (at most kTreeMaxVars), exactly as the reference walks its boxed formulas:
//   Conjunction  propagate = children in order, stop at the first failure (conjunction.rs:96-105);
//                is_subsumed = Kleene and (conjunction.rs:77-94)
//   Disjunction  propagate = read every child's is_subsumed: one True -> nothing to do; all False ->
//                fail; all but one False -> propagate the remaining child (disjunction.rs:96-116);
//                is_subsumed = Kleene or (disjunction.rs:77-94)
//   Boolean b    is_subsumed: b assigned -> (b == 1), else Unknown; propagate: b <- 1
//                (boolean.rs:108-135); BooleanNeg the mirror image (boolean_neg.rs:77-98)
//   leaves       the cmp propagators of pcp_eval.cuh on the private copy
// Both walks are iterative over the prefix node array (no recursion in the kernel): subsumption is
// a reverse scan with a value stack, propagation a work list of node indices.
// ---------------------------------------------------------------------------------------
struct TreeCtx {
  const int* nodes;   // the tree's node array
  IV d[kTreeMaxVars];
};
__device__ __forceinline__ IV tree_rd(const TreeCtx& t, const int* nd, int i) {
  const int slot = nd[3 + 2 * i], off = nd[4 + 2 * i];
  if (slot < 0) return IV{off, off};
  return IV{t.d[slot].lo + off, t.d[slot].hi + off};
}
// narrow operand i of a leaf to [nlo, nhi] (view space); false = empty / Constant outside
__device__ __forceinline__ bool tree_wr(TreeCtx& t, const int* nd, int i, IV cur, int nlo, int nhi) {
  nlo = max(nlo, cur.lo);
  nhi = min(nhi, cur.hi);
  if (nlo > nhi) return false;
  const int slot = nd[3 + 2 * i], off = nd[4 + 2 * i];
  if (slot >= 0) { t.d[slot].lo = nlo - off; t.d[slot].hi = nhi - off; }
  return true;
}
__device__ __noinline__ int tree_leaf_subsumed(const TreeCtx& t, const int* nd) {
  const int type = nd[0];
  if (type == TN_BOOL || type == TN_BOOL_NEG) {
    const IV b = tree_rd(t, nd, 0);
    if (b.lo != b.hi) return 0;
    return ((b.lo == 1) == (type == TN_BOOL)) ? 1 : -1;
  }
  if (type >= TN_LEAF3) {
    const IV x = tree_rd(t, nd, 0), y = tree_rd(t, nd, 1), z = tree_rd(t, nd, 2);
    const int k = type - TN_LEAF3;
    return k == T_GREATER ? sub_greater(x, y, z, 1) : (k == T_LESS ? sub_less(x, y, z, 1) : sub_eq(x, y, z));
  }
  const IV x = tree_rd(t, nd, 0), y = tree_rd(t, nd, 1);
  const int k = type - TN_LEAF;
  if (k == B_LESS) return x.lo >= y.hi ? -1 : (x.hi < y.lo ? 1 : 0);
  const bool same_singleton = x.lo == y.hi && x.hi == y.lo;
  const bool disjoint = x.hi < y.lo || y.hi < x.lo;
  if (k == B_NEQ) return same_singleton ? -1 : (disjoint ? 1 : 0);
  return same_singleton ? 1 : (disjoint ? -1 : 0);
}
__device__ __noinline__ bool tree_leaf_propagate(TreeCtx& t, const int* nd) {
  const int type = nd[0];
  if (type == TN_BOOL || type == TN_BOOL_NEG) {
    const IV b = tree_rd(t, nd, 0);
    const int v = type == TN_BOOL ? 1 : 0;
    return tree_wr(t, nd, 0, b, v, v);  // (outside the domain: the reference's update panics; here the node fails)
  }
  if (type >= TN_LEAF3) {
    IV x = tree_rd(t, nd, 0), y = tree_rd(t, nd, 1), z = tree_rd(t, nd, 2);
    const IV x0 = x, y0 = y, z0 = z;
    const int k = type - TN_LEAF3;
    bool ok;
    if (k == T_GREATER) ok = prop_greater(x, y, z, 1);
    else if (k == T_LESS) ok = prop_less(x, y, z, 1);
    else ok = prop_greater(x, y, z, 0) && prop_less(x, y, z, 0);
    if (!ok) return false;
    return tree_wr(t, nd, 0, x0, x.lo, x.hi) && tree_wr(t, nd, 1, y0, y.lo, y.hi) && tree_wr(t, nd, 2, z0, z.lo, z.hi);
  }
  const IV x = tree_rd(t, nd, 0), y = tree_rd(t, nd, 1);
  const int k = type - TN_LEAF;
  if (k == B_LESS)  // x_less_y.rs:101-108: x first, then y against the old x.lo
    return tree_wr(t, nd, 0, x, x.lo, y.hi - 1) && tree_wr(t, nd, 1, y, x.lo + 1, y.hi);
  if (k == B_EQ) {  // x_eq_y.rs:102-107
    const int lo = max(x.lo, y.lo), hi = min(x.hi, y.hi);
    return tree_wr(t, nd, 0, x, lo, hi) && tree_wr(t, nd, 1, y, lo, hi);
  }
  // XNeqY on Interval: only a value on a bound goes (x_neq_y.rs:82-93)
  if (x.lo == x.hi) {
    if (y.lo == x.lo) return tree_wr(t, nd, 1, y, y.lo + 1, y.hi);
    if (y.hi == x.lo) return tree_wr(t, nd, 1, y, y.lo, y.hi - 1);
  } else if (y.lo == y.hi) {
    if (x.lo == y.lo) return tree_wr(t, nd, 0, x, x.lo + 1, x.hi);
    if (x.hi == y.lo) return tree_wr(t, nd, 0, x, x.lo, x.hi - 1);
  }
  return true;
}
// is_subsumed of the subtree rooted at node r: reverse prefix scan, children's values on a stack
__device__ __noinline__ int tree_subsumed(const TreeCtx& t, int r) {
  signed char val[kTreeMaxNodes];
  int sp = 0;
  const int end = t.nodes[r * kTreeNodeInts + 2];
  for (int i = end - 1; i >= r; --i) {
    const int* nd = t.nodes + i * kTreeNodeInts;
    const int type = nd[0];
    int v;
    if (type == TN_CONJ || type == TN_DISJ) {
      const int n = nd[1];
      v = type == TN_CONJ ? 1 : -1;
      for (int q = 0; q < n; ++q) {
        const int cv = val[--sp];
        v = type == TN_CONJ ? min(v, cv) : max(v, cv);
      }
    } else {
      v = tree_leaf_subsumed(t, nd);
    }
    val[sp++] = (signed char)v;
  }
  return val[0];
}
__device__ __noinline__ bool tree_propagate(TreeCtx& t) {
  short work[kTreeMaxNodes];
  int sp = 0;
  work[sp++] = 0;
  while (sp > 0) {
    const int i = work[--sp];
    const int* nd = t.nodes + i * kTreeNodeInts;
    const int type = nd[0];
    if (type == TN_CONJ) {
      // children in order: push them in reverse (child q starts where child q - 1 ends)
      short kids[kTreeMaxNodes];
      const int n = nd[1];
      int ch = i + 1;
      for (int q = 0; q < n; ++q) { kids[q] = (short)ch; ch = t.nodes[ch * kTreeNodeInts + 2]; }
      for (int q = n - 1; q >= 0; --q) work[sp++] = kids[q];
    } else if (type == TN_DISJ) {
      const int n = nd[1];
      int ch = i + 1, num_disentailed = 0, unknown = -1;
      bool entailed = false;
      for (int q = 0; q < n; ++q) {
        const int k = tree_subsumed(t, ch);
        if (k > 0) { entailed = true; break; }
        if (k < 0) ++num_disentailed; else unknown = ch;
        ch = t.nodes[ch * kTreeNodeInts + 2];
      }
      if (entailed) continue;
      if (num_disentailed == n) return false;
      if (num_disentailed == n - 1) work[sp++] = (short)unknown;
    } else if (!tree_leaf_propagate(t, nd)) {
      return false;
    }
  }
  return true;
}
// One tree propagator (thread 0 of the CTA that got the slot).  Returns 1 if evaluated.
template <bool SMEM>
__device__ __noinline__ unsigned eval_tree(const Params& P, const Ctx& c, int slot, const uint32_t* cur_bits, bool unconditional) {
  if (threadIdx.x != 0) return 0;
  const int b = __ldg(&P.nary_ptr[slot]), k = __ldg(&P.nary_ptr[slot + 1]) - b;
  TreeCtx t;
  t.nodes = P.tree_nodes + __ldg(&P.tree_ptr[slot]);
  IV before[kTreeMaxVars];
  bool any_dirty = unconditional;
  for (int i = 0; i < k; ++i) {
    const int v = __ldg(&P.nary_ops[b + i]).x;
    const int2 d = SMEM ? c.sdom[v] : ldcg_dom(&P.dom[v]);
    t.d[i] = before[i] = IV{d.x, d.y};
    if (!unconditional && ((__ldcg(&cur_bits[v >> 5]) >> (v & 31)) & 1u)) any_dirty = true;
  }
  if (!any_dirty) return 0;
  if (!tree_propagate(t)) { set_failed(c); return 1; }
  const int sub = tree_subsumed(t, 0);
  if (sub < 0) { set_failed(c); return 1; }
  for (int i = 0; i < k; ++i) {
    if (t.d[i].lo == before[i].lo && t.d[i].hi == before[i].hi) continue;
    const int v = __ldg(&P.nary_ops[b + i]).x;
    const int2 cur = ldcg_dom(&P.dom[v]);
    if (!tighten(c, v, 0, IV{cur.x, cur.y}, max(cur.x, t.d[i].lo), min(cur.y, t.d[i].hi))) set_failed(c);
  }
  if (sub > 0) {
    const unsigned bit = 1u << (slot & 31);
    const unsigned old = atomicAnd(&P.nary_active[slot >> 5], ~bit);
    if (old & bit) P.trail[atomicAdd(&P.ctl->trail_cnt, 1u)] = make_ref(F_NARY, (unsigned)slot);
  }
  return 1;
}
#endif

// ---------------------------------------------------------------------------------------
// device-wide barrier; the last CTA to arrive decides whether the fixpoint is reached
// (the "block-reduce of a changed flag": every CTA contributes "I narrowed a variable").
// `decide` = false: plain barrier (after the node prologue).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void node_prologue_finish(const Params& P);
// The arrival word packs three 10-bit counters -- CTAs arrived, CTAs that queued a dirty
// variable, CTAs that saw a failure -- so one release-atomic per CTA carries everything the
// decision needs; the release word `bar_gen` = (generation << kDecBits) | decision, so one
// acquire-load per poll returns both.
__device__ __forceinline__ unsigned grid_barrier(const Params& P, unsigned& gen, unsigned block_props, bool decide,
                                                 const int* s_flags, unsigned iter) {
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
    if ((now & 1023u) == cta_count()) {
      unsigned dec = D_CONTINUE;
      if (decide) {
        if ((now >> 20) & 1023u) dec = D_FAILED;
        else if (((now >> 10) & 1023u) == 0) dec = D_FIXPOINT;
        else if (iter + 1 >= P.max_iterations) dec = D_ITER_CAP;
      }
      if (!decide) node_prologue_finish(P);  // prologue barrier: every CTA has read trail_cnt by now
      else ctl->trail_at_decision = *(volatile unsigned*)&ctl->trail_cnt;
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

// ---------------------------------------------------------------------------------------
// The dirty set: one bit per variable, triple-buffered across iterations (written through
// Ctx::next_bits by apply_updates, read in the next iteration, cleared in the one after).
// At the start of a worklist iteration every CTA compacts the set into the same list in its
// own shared memory (the TMA ring is idle then): ascending variable order, so list index e
// means the same variable everywhere and the rows can be dealt out by index.
// Returns the number of dirty variables; the list is only written when it fits.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int dirty_compact(const uint32_t* bits, int words, int* list, int cap) {
  __shared__ int s_warp[kWarps];
  __shared__ int s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpt = (words + kThreads - 1) / kThreads;  // words per thread (1 for V <= 32768)
  const int w0 = threadIdx.x * wpt, w1 = min(words, w0 + wpt);
  const unsigned m0 = w0 < w1 ? __ldcg(&bits[w0]) : 0u;
  int cnt = __popc(m0);
  for (int w = w0 + 1; w < w1; ++w) cnt += __popc(__ldcg(&bits[w]));
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int x = lane < kWarps ? s_warp[lane] : 0;
    int xi = x;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
    if (lane < kWarps) s_warp[lane] = xi - x;  // exclusive offset of each warp
    if (lane == 31) s_total = xi;
  }
  __syncthreads();
  const int total = s_total;
  if (total <= cap && cnt > 0) {
    int pos = s_warp[warp] + incl - cnt;
    for (int w = w0; w < w1; ++w) {
      unsigned m = w == w0 ? m0 : __ldcg(&bits[w]);
      while (m) { const int b = __ffs(m) - 1; m &= m - 1; list[pos++] = w * 32 + b; }
    }
  }
  __syncthreads();
  return total;
}

// Row-local fixpoint: when the worklist is shorter than the grid, a whole CTA takes one dirty
// variable v and re-evaluates v's row until v itself stops moving.  On bounds-only domains a
// variable typically crawls value by value (each XNeqY against an assigned neighbour trims
// one value off a bound, x_neq_y.rs:82-93): with one device barrier per crawl step a node
// costs tens of iterations; here the crawl runs inside one iteration.  Round 0 gathers the
// row from L2 and leaves its active descriptors in shared memory (the idle TMA ring); the
// later rounds run entirely out of shared memory -- staged descriptors, the CTA's snapshot
// of the domains (kept current by mirroring every update into it) -- with the updates going
// to HBM as fire-and-forget reductions.  Still a chaotic iteration of the same propagators:
// the fixpoint is unchanged.  (A propagator deactivated meanwhile may be evaluated again:
// an entailed propagator prunes nothing, and `deactivate` trails it only once.)
constexpr int kLocalRounds = 1024;
constexpr int kRowCap = 4096;  // staged propagators: 4 B ref + 16 B descriptor word 0 + 4 B unlink slot each
static_assert(kRowStageOff + kRowCap * 24 <= kRingBytes, "row staging area exceeds the ring");
constexpr int kRowBatch = 8;   // row entries a thread keeps in flight in round 0
constexpr int kJumpBits = 1024;  // window of the crawl shortcut, per bound

// The crawl shortcut.  v's row tells which neighbours are assigned: every XNeqY(v, u) over
// plain operands with u a singleton forbids exactly one value of v, and the fixpoint of those
// propagators alone moves lo to the first value >= lo that no assigned neighbour forbids (hi
// symmetrically) -- the composition of the single steps x_neq_y.rs:82-93 would take one by
// one.  While a round evaluates the row, the forbidden values near each bound of v (as it
// stood at the start of the round) are collected in two bit windows in shared memory; after
// the round the bounds jump in one update, and the next round evaluates the row again
// (entailment, failure, the neighbours of a newly assigned v).  Neighbours are read from this
// CTA's snapshot: a stale (wider) view only delays a step to a later iteration.
struct CrawlWin { unsigned bm[2][kJumpBits / 32]; };
__device__ __forceinline__ void crawl_note(const Ctx& c, CrawlWin* w, int v, int2 d, int4 q) {
  const unsigned xv = (unsigned)q.x & kConstVar28;
  if (((unsigned)q.x >> 28) != B_NEQ || xv >= kSumBase28 || q.z < 0) return;
  // X = dom[xv] + q.y must differ from Y = dom[q.z] + q.w
  const bool v_is_x = (int)xv == v;
  const int2 du = c.sdom[v_is_x ? q.z : (int)xv];
  if (du.x != du.y) return;
  const int f = v_is_x ? du.x + q.w - q.y : du.x + q.y - q.w;  // the value of v this neighbour forbids
  if (f >= d.x && f - d.x < kJumpBits) atomicOr(&w->bm[0][(f - d.x) >> 5], 1u << ((f - d.x) & 31));
  if (f <= d.y && d.y - f < kJumpBits) atomicOr(&w->bm[1][(d.y - f) >> 5], 1u << ((d.y - f) & 31));
}
// After the round (every thread of the CTA): move the bounds of v past the forbidden values.
// `d` is the domain the windows are anchored at.  Returns false when every value of the
// domain is forbidden (the reference ends with two equal singletons: failure).
__device__ __forceinline__ bool crawl_jump(const Params& P, Ctx& c, CrawlWin* w, int v, int2 d) {
  __shared__ int2 s_steps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 2) {  // warp 0: steps lo takes upwards; warp 1: steps hi takes downwards
    const unsigned word = w->bm[warp][lane];
    const unsigned open = __ballot_sync(0xffffffffu, word != 0xffffffffu);
    int steps = kJumpBits;
    if (open) {
      const int l = __ffs(open) - 1;
      const unsigned wl = __shfl_sync(0xffffffffu, word, l);
      steps = l * 32 + __ffs(~wl) - 1;
    }
    if (lane == 0) { if (warp == 0) s_steps.x = steps; else s_steps.y = steps; }
  }
  __syncthreads();
  const int up = s_steps.x, down = s_steps.y;
  if ((long long)up + down >= (long long)d.y - d.x + 1) return false;
  if (threadIdx.x == 0 && (up || down)) {
    const int nlo = d.x + up, nhi = d.y - down;
    if (up) { atomicMax(&P.dom[v].x, nlo); c.sdom[v].x = max(c.sdom[v].x, nlo); }
    if (down) { atomicMin(&P.dom[v].y, nhi); c.sdom[v].y = min(c.sdom[v].y, nhi); }
    atomicOr(&c.next_bits[v >> 5], 1u << (v & 31));
    c.flags[0] = 1;
  }
  return true;
}

// Row-local evaluation of an XNeqY over plain variables against the CTA's snapshot: eval_bin's
// B_NEQ branch + finish_eval inline (x_neq_y.rs:82-93, is_subsumed x_neq_y.rs:71-73), every
// narrowing applied to the store (reductions) and mirrored into the snapshot.  Returns true
// when the propagator is entailed: the caller clears its active bit -- with the old bit
// returned, because a propagator shared by two dirty rows can be found entailed twice -- for
// a whole batch at once, so that the round trips of those atomics overlap.
__device__ __forceinline__ bool row_upd(const Params& P, const Ctx& c, int var, int off, IV o, IV n) {
  const bool lo = n.lo > o.lo, hi = n.hi < o.hi;
  if (lo) { atomicMax(&P.dom[var].x, n.lo - off); atomicMax(&c.sdom[var].x, n.lo - off); }
  if (hi) { atomicMin(&P.dom[var].y, n.hi - off); atomicMin(&c.sdom[var].y, n.hi - off); }
  if (lo || hi) atomicOr(&c.next_bits[var >> 5], 1u << (var & 31));
  return lo || hi;
}
__device__ __forceinline__ bool row_eval_neq(const Params& P, const Ctx& c, int4 q) {
  const int xv = (int)((unsigned)q.x & kConstVar28);
  const int2 dx = c.sdom[xv], dy = c.sdom[q.z];
  const IV x{dx.x + q.y, dx.y + q.y}, y{dy.x + q.w, dy.y + q.w};
  if (bin_is_noop(B_NEQ, x, y)) return false;
  IV nx = x, ny = y;
  if (x.lo == x.hi) {
    if (ny.lo == x.lo) ny.lo++; else if (ny.hi == x.lo) ny.hi--;
  } else if (y.lo == y.hi) {
    if (nx.lo == y.lo) nx.lo++; else if (nx.hi == y.lo) nx.hi--;
  }
  if (ny.lo > ny.hi || nx.lo > nx.hi || (nx.lo == ny.hi && nx.hi == ny.lo)) { set_failed(c); return false; }
  bool ch = row_upd(P, c, q.z, q.w, y, ny);
  ch |= row_upd(P, c, xv, q.y, x, nx);
  if (ch) c.flags[0] = 1;
  return nx.hi < ny.lo || ny.hi < nx.lo;
}
__device__ __forceinline__ bool is_plain_neq(int4 q) {
  return ((unsigned)q.x >> 28) == B_NEQ && ((unsigned)q.x & kConstVar28) < kSumBase28 && q.z >= 0;
}
// Append references to the CTA's TrailBuf (warp-aggregated; overflow goes straight to the trail).
__device__ __forceinline__ void tbuf_push(const Params& P, TrailBuf* tb, unsigned ref) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(&tb->n, (unsigned)__popc(m));
  base = __shfl_sync(m, base, leader);
  const unsigned pos = base + __popc(m & lanemask_lt());
  if (pos < (unsigned)kTrailBuf) tb->ref[pos] = ref;
  else P.trail[atomicAdd(&P.ctl->trail_cnt, 1u)] = ref;
}

// One row entry evaluated against the CTA's snapshot (rolled loops only: this code exists once
// per kernel and stays warm in the instruction cache across rounds).  Returns true when the
// inline XNeqY path found the propagator entailed (the caller unlinks it).
template <bool SMEM>
__device__ __forceinline__ bool row_eval_entry(const Params& P, Ctx& c, unsigned ref, int4 q) {
  const unsigned fam = ref >> 29;
  const int slot = (int)(ref & kSlotMask);
  if (fam == F_BIN) {
#ifndef PCP_SET  // (IntervalSet: the staged path does the set work; no crawling on sets)
    if (SMEM && is_plain_neq(q)) return row_eval_neq(P, c, q);
#endif
    eval_loaded<SMEM>(c, F_BIN, slot, q, make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0));
  } else {
    eval_ref<SMEM>(c, fam, slot);
  }
  return false;
}

template <bool SMEM>
__device__ __forceinline__ unsigned expand_rows_local(const Params& P, Ctx& c, const int* list, int n_dirty, unsigned cur_epoch, char* ring, TrailBuf* tb) {
  __shared__ int2 s_before;
  __shared__ int s_nstage, s_nunlink, s_moved;
  __shared__ CrawlWin s_win;
  // staging area (the idle TMA ring behind the dirty list): references, descriptor word 0, and
  // the entries a round found entailed
  unsigned* s_ref = reinterpret_cast<unsigned*>(ring + kRowStageOff);
  int4* s_q0 = reinterpret_cast<int4*>(ring + kRowStageOff + kRowCap * 4);
  unsigned* s_unlink = reinterpret_cast<unsigned*>(ring + kRowStageOff + kRowCap * 20);
  const int lane = threadIdx.x & 31;
  unsigned nprop = 0;
  if (threadIdx.x == 0) c.mirror = SMEM;  // (ordered by the barrier that opens every row)
  for (int e = cta_rank(); e < n_dirty; e += cta_count()) {
    const int v = list[e];
    const int rb = __ldg(&P.adj_ptr[v]), re = __ldg(&P.adj_ptr[v + 1]);
    const bool staged = re - rb <= kRowCap;
    if (threadIdx.x == 0) { s_before = SMEM ? c.sdom[v] : ldcg_dom(&P.dom[v]); s_nstage = 0; s_nunlink = 0; }
    __syncthreads();
    trace_seq(P, 100);
    for (int round = 0;; ++round) {
      const int2 d0 = s_before;                   // v at the start of the round
#ifdef PCP_SET
      const bool crawl = false;                   // a set loses a value wherever it sits: nothing crawls
#else
      const bool crawl = SMEM && d0.x < d0.y;     // an unassigned v can crawl
#endif
      if (crawl && threadIdx.x < 2 * (kJumpBits / 32)) (&s_win.bm[0][0])[threadIdx.x] = 0u;
      if (round == 0 && staged) {
        // Round 0 of a row that fits the staging area: gather it from L2 in batches -- all
        // references of a batch first, then their active words and descriptors: two round
        // trips per batch instead of two per entry -- and leave the active entries in shared
        // memory; the evaluation loop below is the one every later round runs as well.
        for (int j0 = rb; j0 < re; j0 += blockDim.x * kRowBatch) {
          unsigned ref[kRowBatch], word[kRowBatch];
          int4 q0[kRowBatch];
#pragma unroll
          for (int u = 0; u < kRowBatch; ++u) {
            const int j = j0 + u * (int)blockDim.x + (int)threadIdx.x;
            ref[u] = j < re ? __ldg(&P.adj[j]) : 0xffffffffu;
          }
#pragma unroll
          for (int u = 0; u < kRowBatch; ++u) {
            word[u] = 0u;
            q0[u] = make_int4(0, 0, 0, 0);
            if (ref[u] == 0xffffffffu) continue;
            const unsigned fam = ref[u] >> 29;
            const int slot = (int)(ref[u] & kSlotMask);
            const Family& f = P.fam[fam];
            if (slot >= f.n_static) continue;  // truncated by a restore (store.rs:320)
            word[u] = (__ldcg(&f.active[slot >> 5]) >> (slot & 31)) & 1u;
            q0[u] = __ldg(&f.desc[fam == F_DJ ? 3 * (size_t)slot : (size_t)slot]);
          }
#pragma unroll
          for (int u = 0; u < kRowBatch; ++u) {
            const bool keep = word[u] != 0u;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (m) {
              int base = 0;
              if (lane == 0) base = atomicAdd(&s_nstage, __popc(m));
              base = __shfl_sync(0xffffffffu, base, 0);
              if (keep) { const int idx = base + __popc(m & lanemask_lt()); s_ref[idx] = ref[u]; s_q0[idx] = q0[u]; }
            }
          }
        }
      }
      __syncthreads();
      trace_seq(P, 150 + round);
      if (staged) {
        const int n = s_nstage;
#pragma unroll 1
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
          const unsigned ref = s_ref[idx];
          if (ref == 0xffffffffu) continue;  // found entailed in an earlier round
          const int4 q = s_q0[idx];
          if (crawl && (ref >> 29) == F_BIN) crawl_note(c, &s_win, v, d0, q);
          // (no epoch stamp: a propagator shared by two dirty rows may run twice in an
          // iteration, which changes nothing but the count)
          if (row_eval_entry<SMEM>(P, c, ref, q)) {
            s_ref[idx] = 0xffffffffu;
            s_unlink[atomicAdd(&s_nunlink, 1)] = ref;
          }
          ++nprop;
        }
      } else {
        // a row longer than the staging area: gathered again in every round
#pragma unroll 1
        for (int j = rb + (int)threadIdx.x; j < re; j += (int)blockDim.x) {
          const unsigned ref = __ldg(&P.adj[j]);
          const unsigned fam = ref >> 29;
          const int slot = (int)(ref & kSlotMask);
          const Family& f = P.fam[fam];
          if (slot >= f.n_static) continue;
          const unsigned word = __ldcg(&f.active[slot >> 5]);
          const int4 q = __ldg(&f.desc[fam == F_DJ ? 3 * (size_t)slot : (size_t)slot]);
          if (!((word >> (slot & 31)) & 1u)) continue;
          if (crawl && fam == F_BIN) crawl_note(c, &s_win, v, d0, q);
          if (row_eval_entry<SMEM>(P, c, ref, q)) {
            const unsigned bit = 1u << (slot & 31);
            if (atomicAnd(&P.fam[F_BIN].active[slot >> 5], ~bit) & bit) tbuf_push(P, tb, ref);
          }
          ++nprop;
        }
      }
      __syncthreads();
      trace_seq(P, 200 + round);
      if (staged && s_nunlink > 0) {
        // unlink what this round found entailed (store.rs:200-207).  The old bit comes back
        // (a propagator shared by two dirty rows can be found entailed by both); four
        // independent atomics per thread and trip keep their round trips overlapped.
        const int n = s_nunlink;
        for (int i0 = threadIdx.x; i0 < n; i0 += blockDim.x * 4) {
          unsigned r4[4], old[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * (int)blockDim.x;
            r4[u] = i < n ? s_unlink[i] : 0xffffffffu;
            old[u] = 0u;
            if (r4[u] != 0xffffffffu) {
              const int slot = (int)(r4[u] & kSlotMask);
              old[u] = atomicAnd(&P.fam[F_BIN].active[slot >> 5], ~(1u << (slot & 31))) & (1u << (slot & 31));
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (old[u]) tbuf_push(P, tb, r4[u]);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_nunlink = 0;
      }
      if (crawl) {
        if (!crawl_jump(P, c, &s_win, v, d0)) {  // uniform across the CTA
          if (threadIdx.x == 0) set_failed(c);
          break;
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        const int2 a = SMEM ? c.sdom[v] : ldcg_dom(&P.dom[v]);
        s_moved = (a.x != s_before.x || a.y != s_before.y) && a.x <= a.y;
        s_before = a;
      }
      __syncthreads();
      trace_seq(P, 300 + round);
      if (!s_moved || round + 1 >= kLocalRounds) break;
    }
  }
  if (threadIdx.x == 0) c.mirror = false;
  return nprop;
}

template <bool SMEM>
__device__ __forceinline__ unsigned expand_dirty_rows(const Params& P, Ctx& c, const int* list, int n_dirty, unsigned cur_epoch, char* ring, TrailBuf* tb) {
  if (n_dirty <= (int)cta_count()) return expand_rows_local<SMEM>(P, c, list, n_dirty, cur_epoch, ring, tb);
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)cta_rank() * kWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)cta_count() * kWarps;
  const long long S = max(1LL, nwarps / n_dirty);  // segments per row (upper bound)
  const long long items = (long long)n_dirty * S;
  unsigned nprop = 0;
  for (long long item = warp; item < items; item += nwarps) {
    int e = (int)(item / S);
    long long s = item % S;
    int v = list[e];
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
#ifdef PCP_SET
  if (P.restore_from) {  // the label slot holds the bit sets behind the bounds
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.restore_from + P.V);
    const long long nw = (long long)P.V * P.bits_W;
    for (long long i = tid; i < nw; i += nth) P.bits[i] = src[i];
  }
#endif
  if (P.do_trail) {
    unsigned cnt = *(volatile unsigned*)&P.ctl->trail_cnt;
    if (P.trail_keep == 0 && cnt > 4096u) {
      // back to a state in which nothing was entailed (the root of a search, a restart) over a
      // long trail: every allocated propagator is active again -- whole words instead of one
      // reduction per trail entry
      for (int f = 0; f < 4; ++f) {
        uint32_t* act = f == F_NARY ? P.nary_active_w : P.fam[f].active;
        const int n = f == F_NARY ? P.n_nary : P.fam[f].n;
        for (int w = tid; w < (n + 31) / 32; w += nth) {
          const int bits = min(32, n - w * 32);
          const unsigned m = bits == 32 ? 0xffffffffu : ((1u << bits) - 1u);
          if (bits == 32) act[w] = m; else atomicOr(&act[w], m);
        }
      }
      cnt = 0;  // (the loop below has nothing left to do)
    }
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
    // (freshly posted propagators are tail slots)
    if (ip.fam == F_BIN) f.tdesc[ip.slot] = ip.q[0];
    else if (ip.fam == F_TER) { f.tdesc[ip.slot] = ip.q[0]; f.tdescB[ip.slot] = make_int2(ip.q[1].x, ip.q[1].y); }
    else { f.tdesc[3 * (size_t)ip.slot] = ip.q[0]; f.tdesc[3 * (size_t)ip.slot + 1] = ip.q[1]; f.tdesc[3 * (size_t)ip.slot + 2] = ip.q[2]; }
  }
}
__device__ __forceinline__ void node_prologue_finish(const Params& P) {
  if (P.do_trail) P.ctl->trail_cnt = P.trail_keep;
}

// ---------------------------------------------------------------------------------------
// Per-CTA state shared by the single-node kernel and the search-burst kernel.
// ---------------------------------------------------------------------------------------
struct CtaState {
  char* ring;
  int2* sdom;
  uint64_t* full;
  uint64_t* empty;
  int* flags;
  TrailBuf* tbuf;
  uint32_t* dbm;     // dirty bitmap of this CTA's sweeps (dynamic shared memory) or nullptr
  // this CTA's share of a sweep: chunk G of the concatenated families belongs to worker
  // G % workers; per family the first chunk index, and the number of chunks
  int workers, wid, my_chunks;
  int fam_g0[3], fam_cnt[3];
  int pipe_pos;      // chunks this CTA has pushed through the ring so far (all sweeps, all nodes)
  unsigned gen;      // barrier generation
  unsigned rot;      // rotation of the three dirty sets: iteration i of a node works on sets (rot + i) % 3 ...
};

__device__ __forceinline__ FamSweep fam_sweep(const Params& P, const CtaState& st, int fam, int seq_off, int pre) {
  FamSweep a;
  a.desc = P.fam[fam].desc;
  a.descB = P.fam[fam].descB;
  a.cdesc = fam == F_BIN ? P.fam[fam].cdesc : nullptr;
  a.active = P.fam[fam].active;
  a.dom = P.dom;
  a.n = P.fam[fam].n_static;
  a.g0 = st.fam_g0[fam];
  a.workers = st.workers;
  a.cnt = st.fam_cnt[fam];
  a.pipe_pos = st.pipe_pos + seq_off;
  a.first = max(0, min(a.cnt, pre - seq_off));
  a.ring_s = smem_u32(st.ring);
  a.full_s = smem_u32(st.full);
  a.empty_s = smem_u32(st.empty);
  a.sdom_s = st.sdom ? smem_u32(st.sdom) : 0u;
  a.have_aw = 0;
  return a;
}

__device__ __forceinline__ void cta_init(const Params& P, CtaState& st, char* smem, uint64_t* s_full,
                                         uint64_t* s_empty, int* s_flags, bool smem_dom) {
  st.ring = smem;
  st.sdom = smem_dom ? reinterpret_cast<int2*>(smem + kRingBytes) : nullptr;
  st.dbm = P.dirty_bm_off ? reinterpret_cast<uint32_t*>(smem + P.dirty_bm_off) : nullptr;
  st.full = s_full;
  st.empty = s_empty;
  st.flags = s_flags;
  // CTA 0 keeps the books (prologue, posted + tail propagators, result); in a large group the sweep
  // is shared by the other CTAs so that nobody waits for it at the barrier.  In a small group (the
  // contexts of a batched launch: 6 CTAs each) a CTA's share of the sweep is tens of microseconds and
  // the books are a few: there CTA 0 sweeps like everybody else instead of idling a sixth of the SMs.
  const ChunkMap m = chunk_map(P);
  const bool everybody = cta_count() <= (unsigned)kSweepAllCtas;
  st.workers = cta_count() > 1 && !everybody ? (int)cta_count() - 1 : (int)cta_count();
  st.wid = cta_count() > 1 && !everybody ? (int)cta_rank() - 1 : (int)cta_rank();
  st.my_chunks = 0;
  int off = 0;
  for (int f = 0; f < 3; ++f) {
    // my chunks of family f: G = wid + i * workers with off <= G < off + nch[f]
    int i_lo = 0, i_hi = 0;
    if (st.wid >= 0) {
      i_lo = off > st.wid ? (off - st.wid + st.workers - 1) / st.workers : 0;
      const int end = off + m.nch[f];
      i_hi = end > st.wid ? (end - st.wid + st.workers - 1) / st.workers : 0;
    }
    st.fam_cnt[f] = max(0, i_hi - i_lo);
    st.fam_g0[f] = st.wid + i_lo * st.workers - off;
    st.my_chunks += st.fam_cnt[f];
    off += m.nch[f];
  }
  st.pipe_pos = 0;
  st.rot = 0u;
}

// Producer: issue the first chunks of the coming sweep (thread 0 only).  Stages that still
// hold an unconsumed chunk cannot exist here: a sweep is always consumed completely.
__device__ __forceinline__ void pre_issue(const Params& P, CtaState& st) {
  // the ring memory doubles as worklist / n-ary staging (generic-proxy writes): order them
  // before the async-proxy writes of the bulk copies
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int n = min(st.my_chunks, kStages);
  int seq = 0;
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    FamSweep a = fam_sweep(P, st, f, seq, 0);
    for (int i = 0; i < a.cnt && seq + i < n; ++i) {
      const int q = a.pipe_pos + i, s = q % kStages;
      if (q >= kStages) mbar_wait_s(a.empty_s + 8u * s, ((q / kStages) - 1) & 1);
      if (f == 0) producer_issue<0>(a, a.g0 + i * a.workers, a.ring_s + (unsigned)(s * kStageBytes), a.full_s + 8u * s);
      else if (f == 1) producer_issue<1>(a, a.g0 + i * a.workers, a.ring_s + (unsigned)(s * kStageBytes), a.full_s + 8u * s);
      else producer_issue<2>(a, a.g0 + i * a.workers, a.ring_s + (unsigned)(s * kStageBytes), a.full_s + 8u * s);
    }
    seq += a.cnt;
  }
}

// ---------------------------------------------------------------------------------------
// The cold sections of a node -- worklist rows, n-ary propagators, the tail, the prologue, the
// device search's host step -- are compiled out of line, each with its own register
// allocation: inlined into the one big kernel body they cost the hot path (snapshot, posted
// constraint, sweep, barrier) registers, spills and instruction-cache footprint (every feature
// added there measurably slowed the C2 node).  They read the launch parameters from a copy in
// shared memory (`PS`; a pointer to the kernel's own parameter space would make every access
// a generic load).
// ---------------------------------------------------------------------------------------
template <bool SMEM>
__device__ __noinline__ unsigned rows_outlined(const Params* PS, Ctx* c, const int* list, int n_dirty, unsigned cur_epoch,
                                               char* ring, TrailBuf* tb) {
  return expand_dirty_rows<SMEM>(*PS, *c, list, n_dirty, cur_epoch, ring, tb);
}
template <bool SMEM>
__device__ __noinline__ unsigned nary_outlined(const Params* PS, const Ctx* c, char* ring, const uint32_t* cur_bits, bool all) {
  const Params& P = *PS;
  unsigned n = 0;
  // dealt from the last CTA backwards: CTA 0 (prologue, posted and tail propagators) is the
  // last to get one
  for (int s = (int)cta_count() - 1 - (int)cta_rank(); s < P.n_nary; s += cta_count()) {
    if (!((__ldcg(&P.nary_active[s >> 5]) >> (s & 31)) & 1u)) continue;
    const int kind = __ldg(&P.nary_kind[s]);
    unsigned ev;
#if !defined(PCP_BIN_ONLY) && !defined(PCP_SET)
    if (kind == N_TREE) ev = eval_tree<SMEM>(P, *c, s, cur_bits, all);
    else
#endif
    if (kind == N_ALL_EQUAL) ev = eval_all_equal<SMEM>(P, *c, s, cur_bits, all);
#ifdef PCP_SET
    else ev = eval_distinct_set<SMEM>(P, *c, s, ring, cur_bits, all);
#else
    else ev = eval_distinct<SMEM>(P, *c, s, ring, cur_bits, all);
#endif
    if (threadIdx.x == 0) n += ev;
  }
  return n;
}
// older tail propagators: CTA 0, every iteration (the posted ones are handled by the caller),
// after the refresh so that they read this CTA's snapshot: active word and descriptor in one
// round trip, no gather of the domains
template <bool SMEM>
__device__ __noinline__ unsigned tail_outlined(const Params* PS, const Ctx* c, int bin_n, bool first_iter, int n_inline,
                                               const InlineProp* inl) {
  const Params& P = *PS;
  unsigned n = 0;
  for (unsigned fam = 0; fam < 3; ++fam) {
    const Family& f = P.fam[fam];
    const int fn = fam == F_BIN ? bin_n : f.n;
    for (int p = f.n_static + threadIdx.x; p < fn; p += blockDim.x) {
      bool is_inl = false;
      if (first_iter)
        for (int i = 0; i < n_inline; ++i) is_inl |= inl[i].fam == fam && inl[i].slot == p;
      if (is_inl) continue;
      const unsigned word = __ldcg(&f.active[p >> 5]);
      int4 q0, q1, q2;
      load_desc(f, fam, p, q0, q1, q2);
      if ((word >> (p & 31)) & 1u) { eval_loaded<SMEM>(*c, fam, p, q0, q1, q2); ++n; }
    }
  }
  return n;
}
__device__ __noinline__ void prologue_outlined(const Params* PS, int tid, int nth) { node_prologue(*PS, tid, nth); }

#ifdef PCP_SET
// Bring the cached bounds of a narrowed variable back onto elements of its set: two concurrent
// updates can leave lo on a value that a third thread removed meanwhile (the set itself is
// always right: bits are only cleared).  Every CTA does this when it refreshes its view of the
// dirty variables; the result also goes back to the store (idempotent reductions, not a
// narrowing of the domain: nobody is re-scheduled for it).  Returns 1 when the domain is empty.
__device__ __forceinline__ int set_normalise(const Params& P, int v, int2& d) {
  if (d.x > d.y) return 1;
  const SetW sw = set_of(P);
  const int lo = sb_next(sw, v, d.x, d.y);
  if (lo == INT32_MAX) { d.x = d.y + 1; return 1; }
  const int hi = sb_prev(sw, v, d.y, lo);
  if (lo != d.x) atomicMax(&P.dom[v].x, lo);
  if (hi != d.y) atomicMin(&P.dom[v].y, hi);
  d.x = lo;
  d.y = hi;
  return 0;
}
#endif

// One node's fixpoint: iteration 0 (posted propagators, tail, streaming sweep or seeded
// worklist) and the worklist / re-sweep iterations, each closed by the deciding barrier.
// Every thread of every CTA calls it with the same arguments; returns the decision, `iters`
// the number of iterations.  `bin_n` is the end of the binary tail (dynamic during a burst).
// On return the three dirty sets are empty again.
template <bool SMEM>
__device__ __forceinline__ unsigned fixpoint_node(const Params& P, const Params* PS, CtaState& st, unsigned epoch0, int bin_n,
                                                  int n_inline, const InlineProp* inl, bool full_sweep,
                                                  int seeded, bool pre_issued, unsigned& iters_out) {
  __shared__ unsigned s_wprops[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = P.dirty_words;
  int* const list = reinterpret_cast<int*>(st.ring);
  // The evaluation context is uniform across the CTA and lives in shared memory: as a local
  // variable handed to the out-of-line evaluators by reference it sat in local memory, and
  // every reload of one of its fields could miss L1 (the CTA's shared memory leaves little
  // of it) and cost an L2 round trip on the critical path.  Thread 0 writes, a barrier orders.
  __shared__ Ctx s_ctx;
  __shared__ TrailBuf s_tbuf;
  Ctx& c = s_ctx;
  st.tbuf = &s_tbuf;
  if (threadIdx.x == 0) {
    s_tbuf.n = 0u;
    c.P = &P;
    c.sdom = st.sdom;
    c.sdom_s = SMEM ? smem_u32(st.sdom) : 0u;
    c.flags = st.flags;
    c.mirror = false;
    c.dom_w = P.dom;
    c.trail = P.trail;
    c.trail_cnt = &P.ctl->trail_cnt;
    c.flags_s = smem_u32(st.flags);
    c.tbuf_s = smem_u32(st.tbuf);
    c.dbm_s = st.dbm ? smem_u32(st.dbm) : 0u;
  }
  ActiveWords aw;
  aw.w[0] = aw.w[1] = aw.w[2] = aw.w[3] = 0u;
  uint4 awc0 = make_uint4(0u, 0u, 0u, 0u), awc1 = awc0;  // compact binary stream: eight words
  const bool compact = P.fam[0].cdesc != nullptr;
  // the active words of this CTA's first chunk: prefetched while the prologue settles
  const int fam_first = st.fam_cnt[0] > 0 ? 0 : (st.fam_cnt[1] > 0 ? 1 : 2);
  if (warp > 0 && st.my_chunks > 0 && full_sweep && fam_first < 2) {
    if (fam_first == 0 && compact) {
      const int base = st.fam_g0[0] * kChunkBinC + (warp - 1) * 32 * kGroupsBinC;
      if (base < P.fam[0].n_static) {
        const uint4* p = reinterpret_cast<const uint4*>(P.fam[0].active + (base >> 5));
        awc0 = __ldcg(p);
        awc1 = __ldcg(p + 1);
      }
    } else if (fam_first == 0) {
      aw = load_active<kGroupsBin>(P.fam[0].active, st.fam_g0[0] * kChunkBin + (warp - 1) * 32 * kGroupsBin, P.fam[0].n_static);
    } else {
      aw = load_active<kGroupsTer>(P.fam[1].active, st.fam_g0[1] * kChunkTer + (warp - 1) * 32 * kGroupsTer, P.fam[1].n_static);
    }
  }
  unsigned iter = 0, dec, nprop = 0;
  int cur_buf = 0, next_buf = 1;
  while (true) {
    cur_buf = (int)((st.rot + iter) % 3u);
    next_buf = (int)((st.rot + iter + 1u) % 3u);
    const int spare_buf = (int)((st.rot + iter + 2u) % 3u);
    const unsigned cur_epoch = epoch0 + iter;
    const uint32_t* cur_bits = P.dirty_bits + (size_t)cur_buf * W;
    if (threadIdx.x == 0) {
      c.next_bits = P.dirty_bits + (size_t)next_buf * W;
      c.local = false;
      c.mark_dirty = true;
      c.bookkeep = true;
    }
    __syncthreads();
    trace_mark1(P, iter, 0);
    // the spare set was read in the previous iteration (of this node, or -- when the device search
    // went straight on to the next node -- the last one of the previous node) and is written in the
    // next one: cleared here, before this iteration's barrier
    if (cta_rank() == 0)
      for (int w = threadIdx.x; w < W; w += blockDim.x) P.dirty_bits[(size_t)spare_buf * W + w] = 0u;

    // Incremental launch (no schedule-everything sweep): the variables of the posted propagators go
    // straight into the worklist of THIS iteration -- every CTA has applied the posted propagators to
    // its own view, so the rows can be expanded at once instead of after a first barrier (one
    // iteration per node instead of two when nothing else moves).  Only for posted binary
    // propagators over plain / constant operands in a store without n-ary propagators (those find
    // their dirty operands in the shared bit set).
    bool posted_rows = false;
    if (iter == 0 && !full_sweep && !seeded && n_inline > 0 && P.n_nary == 0) {
      posted_rows = true;
      for (int i = 0; i < n_inline; ++i)
        posted_rows = posted_rows && inl[i].fam == F_BIN && dec_var28((unsigned)inl[i].q[0].x) >= -1 && inl[i].q[0].z >= -1;
#ifdef PCP_SET
      for (int i = 0; i < n_inline; ++i) posted_rows = posted_rows && ((unsigned)inl[i].q[0].x >> 28) == B_LESS;
#endif
    }
    if (iter == 0 && n_inline > 0) {
      // Propagators posted since the last node.  With a full sweep ahead they need not enter
      // the worklist: every CTA applies them to its own view of the domains first, so the sweep
      // already sees their effect.
      if (threadIdx.x == 0) {
        c.mark_dirty = !full_sweep && !posted_rows;
#ifdef PCP_SET
        // A posted XLessY (what BinarySplit and BranchAndBound post) only moves bounds, which every
        // CTA applies to its own snapshot; anything else may clear bits of the shared sets behind the
        // back of a CTA that is already sweeping, so what it narrows is queued like any other change.
        for (int i = 0; i < n_inline; ++i)
          if (inl[i].fam != F_BIN || ((unsigned)inl[i].q[0].x >> 28) != B_LESS) c.mark_dirty = true;
#endif
        if (cta_rank() == 0 || !SMEM) {  // the global store: counted and book-kept by CTA 0
          c.bookkeep = cta_rank() == 0;
          for (int i = 0; i < n_inline; ++i)
            eval_full<false>(c, inl[i].fam, inl[i].slot, inl[i].q[0], inl[i].q[1], inl[i].q[2]);
          if (cta_rank() == 0) nprop += n_inline;
          if (!SMEM) __threadfence();
        }
        if (SMEM) {
          c.local = true;
          c.bookkeep = false;
          for (int i = 0; i < n_inline; ++i)
            eval_full<true>(c, inl[i].fam, inl[i].slot, inl[i].q[0], inl[i].q[1], inl[i].q[2]);
        }
        c.local = false;
        c.mark_dirty = true;
        c.bookkeep = true;
      }
      __syncthreads();
    }
    if (iter == 0) trace_mark(P, 2);
    // ---- the variables narrowed in the previous iteration (iteration 0 of an incremental
    // launch: seeded by the host).  Every CTA derives the same count, hence the same choice
    // between expanding their rows and sweeping again.
    int n_dirty = 0;
    bool sweep_now = iter == 0 && full_sweep, skip = false;
    if (iter > 0 || (!full_sweep && (seeded || posted_rows))) {
      n_dirty = (iter > 0 || seeded) ? dirty_compact(cur_bits, W, list, kListCap) : 0;
      const int n_refresh = n_dirty;  // (the posted variables appended below are current in every CTA's own view)
      if (posted_rows && n_dirty + 2 * kMaxInline <= kListCap) {
        __shared__ int s_nd;
        __syncthreads();
        if (threadIdx.x == 0) {
          int n = n_dirty;
          for (int i = 0; i < n_inline; ++i) {
            const int vs[2] = {dec_var28((unsigned)inl[i].q[0].x), inl[i].q[0].z};
            for (int q = 0; q < 2; ++q) {
              if (vs[q] < 0) continue;
              bool have = false;
              for (int j = 0; j < n; ++j) have |= list[j] == vs[q];
              if (!have) list[n++] = vs[q];
            }
          }
          s_nd = n;
        }
        __syncthreads();
        n_dirty = s_nd;
      }
      trace_mark1(P, iter, 1);
      // many dirty variables: their CSR rows cover most of the store, and a streaming sweep is
      // cheaper than gathering the rows
      if (n_dirty > kListCap || (long long)n_dirty * 8 >= (long long)P.V) sweep_now = true;
      // refresh the snapshot and catch domains emptied by two concurrent updates (with a
      // snapshot every CTA needs every dirty variable; without one the check is shared by the
      // grid)
      int bad = 0;
      if (sweep_now) {
        const int r0 = SMEM ? threadIdx.x : cta_rank() * blockDim.x + threadIdx.x;
        const int rs = SMEM ? blockDim.x : cta_count() * blockDim.x;
        for (int v = r0; v < P.V; v += rs) {  // coalesced re-read of everything beats gathers
          int2 d = ldcg_dom(&P.dom[v]);
#ifdef PCP_SET
          bad |= set_normalise(P, v, d);
#endif
          if (SMEM) st.sdom[v] = d;
          bad |= d.x > d.y;
        }
      } else {
        const int r0 = SMEM ? threadIdx.x : cta_rank() * blockDim.x + threadIdx.x;
        const int rs = SMEM ? blockDim.x : cta_count() * blockDim.x;
        for (int i = r0; i < n_refresh; i += rs) {
          int v = list[i];
          int2 d = ldcg_dom(&P.dom[v]);
#ifdef PCP_SET
          bad |= set_normalise(P, v, d);
#endif
          if (SMEM) st.sdom[v] = d;
          bad |= d.x > d.y;
        }
      }
      if (SMEM && posted_rows && sweep_now) {
        // the whole snapshot was just re-read from the store, where CTA 0's application of the posted
        // propagators may not have landed yet: apply them to this CTA's view again
        __syncthreads();
        if (threadIdx.x == 0) {
          c.local = true; c.bookkeep = false; c.mark_dirty = false;
          for (int i = 0; i < n_inline; ++i) eval_full<true>(c, inl[i].fam, inl[i].slot, inl[i].q[0], inl[i].q[1], inl[i].q[2]);
          c.local = false; c.bookkeep = true; c.mark_dirty = true;
        }
        __syncthreads();
      }
      if (__syncthreads_or(bad)) {
        // failed: with a snapshot every CTA sees it and the iteration's work is skipped
        if (threadIdx.x == 0) set_failed(c);
        skip = SMEM;
      }
      trace_mark1(P, iter, 2);
      if (!skip && !sweep_now && n_dirty > 0) nprop += rows_outlined<SMEM>(PS, &c, list, n_dirty, cur_epoch, st.ring, st.tbuf);
    }
    if (cta_rank() == 0 && !skip && (P.fam[0].n_static < bin_n || P.fam[1].n_static < P.fam[1].n || P.fam[2].n_static < P.fam[2].n))
      nprop += tail_outlined<SMEM>(PS, &c, bin_n, iter == 0, n_inline, inl);
    if (sweep_now && !skip && st.my_chunks > 0) {
      // ---- the streaming sweep over the static descriptor arrays (ring positions keep
      // counting across sweeps so the mbarrier phases stay consistent).  (`skip` is never set
      // in iteration 0, so pre-issued chunks are always consumed.)
      const int pre = (iter == 0 && pre_issued) ? min(st.my_chunks, kStages) : 0;
      // the ring memory doubles as worklist / n-ary staging (generic-proxy writes): order them
      // before the async-proxy writes of the next bulk copies
      if (threadIdx.x == 0 && pre == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const bool have_aw = iter == 0 && full_sweep;  // prefetched above
      const uint4 aw4 = make_uint4(aw.w[0], aw.w[1], aw.w[2], aw.w[3]);
      unsigned n = 0;
      int seq = 0;
      if (st.fam_cnt[0] > 0) {
        FamSweep a = fam_sweep(P, st, 0, seq, pre);
        a.have_aw = have_aw && fam_first == 0;
        const bool lean = P.fam[0].all_plain && P.fam[0].kind_mask == (1 << B_NEQ);
        if (compact) n += sweep_bin_compact<SMEM>(c, a, awc0, awc1);
        else n += lean ? sweep_bin<SMEM, true>(c, a, aw4) : sweep_bin<SMEM, false>(c, a, aw4);
        seq += a.cnt;
      }
#ifndef PCP_BIN_ONLY
      if (st.fam_cnt[1] > 0) {
        FamSweep a = fam_sweep(P, st, 1, seq, pre);
        a.have_aw = have_aw && fam_first == 1;
        const bool lean = P.fam[1].all_plain && P.fam[1].kind_mask == (1 << T_EQ);
        n += lean ? sweep_ter<SMEM, true>(c, a, aw4) : sweep_ter<SMEM, false>(c, a, aw4);
        seq += a.cnt;
      }
      if (st.fam_cnt[2] > 0) {
        FamSweep a = fam_sweep(P, st, 2, seq, pre);
        n += sweep_dj<SMEM>(c, a);
        seq += a.cnt;
      }
#endif
      if (lane == 0) nprop += n;  // counted per warp
      st.pipe_pos += st.my_chunks;
      if (st.dbm) {  // what this CTA's sweep narrowed joins the dirty set of the next iteration
        __syncthreads();
        for (int w = threadIdx.x; w < W; w += blockDim.x) {
          const unsigned m = st.dbm[w];
          if (m) { atomicOr(&c.next_bits[w], m); st.dbm[w] = 0u; }
        }
      }
    }
    if (iter <= 1 && P.trace) { __syncthreads(); if (iter == 0) trace_mark(P, 3); else trace_mark1(P, iter, 3); }
    // n-ary propagators: one CTA each; re-run when one of their operands is dirty.
#ifndef PCP_BIN_ONLY
    if (P.n_nary > 0 && !skip && (iter > 0 || full_sweep || n_dirty > 0))
      nprop += nary_outlined<SMEM>(PS, &c, st.ring, cur_bits, iter == 0 && full_sweep);
#endif

    // block-level propagation count (per-warp partial sums in shared memory, no atomics),
    // then the barrier + decision
    for (int o = 16; o; o >>= 1) nprop += __shfl_xor_sync(0xffffffffu, nprop, o);
    if (lane == 0) s_wprops[warp] = nprop;
    nprop = 0;
    __syncthreads();
    trail_flush(P, &s_tbuf);  // the propagators this CTA's sweep found entailed
    unsigned bp = 0;
    if (warp == 0) {
      bp = lane < kWarps ? s_wprops[lane] : 0u;
      for (int o = 16; o; o >>= 1) bp += __shfl_xor_sync(0xffffffffu, bp, o);
    }
    if (iter == 0) trace_mark(P, 4);
    trace_mark1(P, iter, 4);
    dec = grid_barrier(P, st.gen, bp, true, st.flags, iter);
    if (iter == 0) trace_mark(P, 5);
    trace_mark1(P, iter, 5);
    if (P.trace && cta_rank() == 0 && threadIdx.x == 0 && iter < 32) {  // per-iteration record
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.trace[8 * 256 + iter * 4 + 0] = t;
      P.trace[8 * 256 + iter * 4 + 1] = (unsigned long long)n_dirty;
      P.trace[8 * 256 + iter * 4 + 2] = dec | (sweep_now ? 8u : 0u);
    }
    ++iter;
    if (dec != D_CONTINUE) break;
  }
  // The sets this node leaves behind: `cur` (read in the last iteration) still holds the marks of the
  // iteration before, `next` is empty at a fixpoint (nobody narrowed anything) and holds marks after a
  // failure.  The caller either clears them (dirty_sets_reset: single-node launches, and a device
  // search that meets at a barrier before the next node) or -- fast descent, fixpoints only -- lets the
  // rotation run on: `next` becomes the next node's current set, `cur` its spare, cleared by CTA 0 in
  // that node's first iteration before anybody writes it.  (Clearing them here raced with the CTAs that
  // had already started the next node: a late CTA 0 wiped their first marks.)
  st.rot = (st.rot + iter) % 3u;  // = the index of the last iteration's `next`
  iters_out = iter;
  return dec;
}
// every thread of CTA 0; the other CTAs only realign their rotation
__device__ __forceinline__ void dirty_sets_reset(const Params& P, CtaState& st) {
  if (cta_rank() == 0)
    for (int w = threadIdx.x; w < 3 * P.dirty_words; w += blockDim.x) P.dirty_bits[w] = 0u;
  st.rot = 0u;
}

// One node of one engine: prologue (restore, posted propagators), the fixpoint, the epilogue (label
// copy, result header, zero-copy mirror).  `P`: the launch parameters where the hot loops read them
// (the kernel's own parameter space, or shared memory in a batched launch); `PS`: their copy in
// shared memory for the out-of-line cold sections.
template <bool SMEM>
__device__ __forceinline__ void fixpoint_launch_body(const Params& P, const Params* PS, char* smem, uint64_t* s_full,
                                                     uint64_t* s_empty, int* s_flags) {
  CtaState st;
  cta_init(P, st, smem, s_full, s_empty, s_flags, SMEM);
  if (st.dbm) for (int w = threadIdx.x; w < P.dirty_words; w += blockDim.x) st.dbm[w] = 0u;  // (ordered by the barrier below)
  Control* ctl = P.ctl;

  if (threadIdx.x == 0) {
    s_flags[0] = s_flags[1] = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // start streaming descriptors right away: they do not depend on the node prologue
    if (P.full_sweep) pre_issue(P, st);
  }
  // without a restore the domains are already final: snapshot them while the TMA runs
  if (SMEM && !P.sync0)
    for (int v = threadIdx.x; v < P.V; v += blockDim.x) st.sdom[v] = ldcg_dom(&P.dom[v]);
  __syncthreads();
  // barrier generation and epoch travel in the launch parameters (the host tracks both from
  // the result header): no dependent global load before the first useful one
  st.gen = P.gen0;
  const unsigned epoch0 = P.epoch0;
  trace_mark(P, 0);

  if (P.sync0) {
    prologue_outlined(PS, cta_rank() * blockDim.x + threadIdx.x, cta_count() * blockDim.x);
    grid_barrier(P, st.gen, 0, false, s_flags, 0);  // its last arriver resets trail_cnt (nobody pushes yet)
    if (SMEM) {
      for (int v = threadIdx.x; v < P.V; v += blockDim.x) st.sdom[v] = ldcg_dom(&P.dom[v]);
      __syncthreads();
    }
  } else if (cta_rank() == 0) {
    prologue_outlined(PS, threadIdx.x, blockDim.x);  // CTA-local effects only (posted propagators)
    __syncthreads();
  }
  trace_mark(P, 1);

  unsigned iters = 0;
  const unsigned dec = fixpoint_node<SMEM>(P, PS, st, epoch0, P.fam[F_BIN].n, P.n_inline, P.inl, P.full_sweep != 0,
                                           P.seed_dirty, P.full_sweep != 0, iters);

  dirty_sets_reset(P, st);  // the sets are empty between launches
  trace_mark(P, 6);
#ifdef PCP_SET
  // the label copy of the bit sets: every CTA takes a slice (at a fixpoint nobody writes them any more)
  if (P.snapshot_to && dec == D_FIXPOINT) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(P.snapshot_to + P.V);
    const long long nw = (long long)P.V * P.bits_W;
    for (long long i = (long long)cta_rank() * blockDim.x + threadIdx.x; i < nw; i += (long long)cta_count() * blockDim.x)
      dst[i] = __ldcg(&P.bits[i]);
  }
#endif
  if (cta_rank() == 0) {
    // Branch::distribute takes a label right after an Unknown node (branch.rs:36-49): leave the
    // copy of the fixpoint domains in the next label slot so that pcp_label is free.
    // (the last iteration of a fixpoint narrowed nothing anywhere, so this CTA's snapshot --
    // refreshed at the start of that iteration -- is the store: no reload)
    if (P.snapshot_to && dec == D_FIXPOINT)
      for (int v = threadIdx.x; v < P.V; v += blockDim.x) P.snapshot_to[v] = SMEM ? st.sdom[v] : ldcg_dom(&P.dom[v]);
    if (P.host_result && P.host_dom) {  // the domains go straight to the host's mirror
      int2* hd = reinterpret_cast<int2*>(P.host_result + 1);
      const bool from_snapshot = SMEM && dec == D_FIXPOINT;
      for (int v = threadIdx.x; v < P.V; v += blockDim.x) hd[v] = from_snapshot ? st.sdom[v] : ldcg_dom(&P.dom[v]);
#ifdef PCP_SET
      if (P.host_sizes && dec != D_FAILED) {  // Cardinality::size() per variable, for the host's FirstSmallestVar
        const SetW sw = set_of(P);
        for (int v = threadIdx.x; v < P.V; v += blockDim.x) {
          const int2 d = from_snapshot ? st.sdom[v] : ldcg_dom(&P.dom[v]);
          P.host_sizes[v] = d.x > d.y ? 0u : sb_count(sw, v, d.x, d.y);
        }
      }
#endif
      __threadfence_system();
    }
    if (P.host_result) __syncthreads();
    if (threadIdx.x == 0) {
      Result r;
      r.failed = dec == D_FAILED;
      r.trail_cnt = *(volatile unsigned*)&ctl->trail_cnt;
      r.iterations = iters;
      r.epoch = epoch0 + iters + 1;
      r.propagations = *(volatile unsigned long long*)&ctl->propagations;
      r.decision = dec;
      r.gen = st.gen;
      r.seq = 0u;
      *P.result = r;
      ctl->epoch = epoch0 + iters + 1;
      ctl->iterations = iters;
      ctl->last_decision = dec;
      if (P.host_result) {
        Result* h = P.host_result;
        h->failed = r.failed; h->trail_cnt = r.trail_cnt; h->iterations = r.iterations; h->epoch = r.epoch;
        h->propagations = r.propagations; h->decision = r.decision; h->gen = r.gen;
        __threadfence_system();
        *(volatile unsigned*)&h->seq = P.host_seq;
      }
    }
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(kThreads, 1) pcp_fixpoint_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
  __shared__ int s_flags[2];
  // layout: [ring | n-ary staging (aliased)] [domain snapshot]
  __shared__ Params s_P;  // copy of the launch parameters for the out-of-line cold sections
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(&P);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_P);
    for (int i = threadIdx.x; i < (int)(sizeof(Params) / 4); i += blockDim.x) dst[i] = src[i];
  }
  cta_place(blockIdx.x, gridDim.x);  // (its barrier also orders the copy above)
  fixpoint_launch_body<SMEM>(P, &s_P, smem, s_full, s_empty, s_flags);
}

// One node of each of several engines in ONE launch (pcp_consistency_batch on engines that run
// the same kernel variant with the same geometry: forks of one model, each in its own subtree).
// CTAs [k * group, (k + 1) * group) serve engine k with the launch parameters batch[k]; every
// engine keeps its own barrier word, counters, dirty sets and result block, so the groups never
// meet.  Same work per engine as pcp_fixpoint_kernel; what the batch removes is the launch per
// engine: K cooperative launches on K streams start about 3 us apart on the device (70 us for
// K = 24, as long as a node of a 6-CTA group itself takes).
__global__ void __launch_bounds__(kThreads, 1) pcp_fixpoint_batch_kernel(const Params* __restrict__ batch, int group) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
  __shared__ int s_flags[2];
  __shared__ Params s_P;
  const unsigned k = blockIdx.x / (unsigned)group;
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(batch + k);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_P);
    for (int i = threadIdx.x; i < (int)(sizeof(Params) / 4); i += blockDim.x) dst[i] = __ldg(&src[i]);
  }
  cta_place(blockIdx.x - k * (unsigned)group, (unsigned)group);
  fixpoint_launch_body<true>(s_P, &s_P, smem, s_full, s_empty, s_flags);
}

// ---------------------------------------------------------------------------------------
// Search burst: the callers of the path -- Brancher(FirstSmallestVar,
// MiddleVal, BinarySplit) under OneSolution/AllSolution/StopNode -- executed on the device
// between fixpoints, so that a run of DFS nodes costs one launch.  CTA 0 plays the host:
// after a node's final barrier it derives the status (propagation/store.rs:250-256), selects
// the branching variable (search/branching/first_smallest_var.rs:30-39: first index among the
// smallest domains of size > 1) and value (middle_val.rs:25-27), takes the label
// (branch.rs:36-49: copy of the domains + trail length + number of propagators), pushes the
// two alternatives (right first, so the left child is explored first,
// engine/one_solution.rs:46-51), pops the next branch, restores its label (branch.rs:51-55)
// and posts its constraint (binary_split.rs:46-57) in the tail; the other CTAs wait at a
// barrier and meanwhile stream the descriptors of the next sweep into shared memory.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void burst_load(const BurstCtl* bc, BurstLocal* L) {
  L->n_branch = bc->n_branch; L->n_labels = bc->n_labels; L->cur_label = bc->cur_label; L->bin_n = bc->bin_n;
  L->stopped = bc->stopped; L->root_pending = bc->root_pending; L->last_status = bc->last_status; L->err = bc->err;
  L->nodes = bc->nodes; L->solutions = bc->solutions; L->failures = bc->failures; L->iterations = bc->iterations;
  L->top_valid = 0;
  L->run = 0;
}
__device__ __forceinline__ void burst_store(BurstCtl* bc, const BurstLocal* L) {
  bc->n_branch = L->n_branch; bc->n_labels = L->n_labels; bc->cur_label = L->cur_label; bc->bin_n = L->bin_n;
  bc->stopped = L->stopped; bc->root_pending = L->root_pending; bc->last_status = L->last_status; bc->err = L->err;
  bc->nodes = L->nodes; bc->solutions = L->solutions; bc->failures = L->failures; bc->iterations = L->iterations;
}

// CTA 0, all threads: the host's work between two fixpoints.  `have_node`: a node just reached
// its decision `dec` (false at the start of a launch).  `scratch`: V int2 of shared memory
// (CTA 0's snapshot area) or nullptr.  Posts the next node through `bc` or stops the burst.
__device__ __noinline__ void burst_host_step(const Params& P, const BurstParams& B, BurstLocal* L, int2* scratch,
                                                bool have_node, unsigned dec, unsigned iters, unsigned long long done) {
  __shared__ unsigned long long s_best[kWarps];
  __shared__ int4 s_pop;
  __shared__ int2 s_pop_meta;
  BurstCtl* bc = B.bc;
  Control* ctl = P.ctl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (have_node && dec == D_ITER_CAP) { if (tid == 0) L->err = 2; have_node = false; }
  if (have_node) {
    // ---- close the node: status, trace, branching, label.  The trail length is the one published with
    // the node's decision: in a fast descent the other CTAs are already running the left child, and what
    // they find entailed there must not count as entailed at this node (its label would then leave those
    // propagators switched off in the right subtree).
    const unsigned trail_cnt = __ldcg(&ctl->trail_at_decision);
    const int bin_n = L->bin_n;
    const unsigned long long n = L->nodes;
    unsigned long long best = ~0ull;
    // one pass over the final domains; at a fixpoint CTA 0's snapshot (`scratch`) already is
    // the store -- the last iteration narrowed nothing -- so nothing is reloaded
    const bool from_snapshot = scratch != nullptr && dec == D_FIXPOINT;
    for (int v = tid; v < P.V; v += blockDim.x) {
      int2 d = from_snapshot ? scratch[v] : ldcg_dom(&P.dom[v]);
      if (scratch && !from_snapshot) scratch[v] = d;
#ifdef PCP_SET
      unsigned size = d.x >= d.y ? (unsigned)(d.y - d.x) + 1u : sb_count(set_of(P), v, d.x, d.y);  // Cardinality::size()
#else
      unsigned size = (unsigned)(d.y - d.x) + 1u;
#endif
      if (size > 1u) best = min(best, ((unsigned long long)size << 32) | (unsigned)v);
    }
    // propagation/store.rs:250-256
    const int status = dec == D_FAILED ? -1 : ((long long)trail_cnt == B.props_base + bin_n ? 1 : 0);
    const bool stop = B.node_limit && n + 1 >= B.node_limit;
    if (scratch) __syncthreads();
    if (n < B.t_cap) {
      if (tid == 0 && B.t_status) B.t_status[n] = status;
      if (B.t_dom && status != -1)
        for (int v = tid; v < P.V; v += blockDim.x)
          B.t_dom[n * (unsigned long long)P.V + v] = scratch ? scratch[v] : ldcg_dom(&P.dom[v]);
#ifdef PCP_SET
      if (B.t_bits && status != -1) {
        const long long nw = (long long)P.V * P.bits_W;
        for (long long i = tid; i < nw; i += blockDim.x) B.t_bits[n * (unsigned long long)nw + i] = __ldcg(&P.bits[i]);
      }
#endif
    }
    if (status == 0 && !stop) {
      // FirstSmallestVar: min over (size, index) of the variables with size > 1
      for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (lane == 0) s_best[warp] = best;
      __syncthreads();
      best = s_best[0];
      for (int w = 1; w < kWarps; ++w) best = min(best, s_best[w]);
      const int var = (int)(best & 0xffffffffu);
      const int slot = L->n_labels, nb = L->n_branch;
      if (best == ~0ull || slot >= B.max_labels || nb + 2 > B.max_branches) {
        __syncthreads();
        if (tid == 0) L->err = 1;
      } else {
        const int2 d = scratch ? scratch[var] : ldcg_dom(&P.dom[var]);
        const int val = (d.x + d.y) / 2;  // MiddleVal: truncating division
        int2* dst = B.stack + (long long)slot * B.stack_stride;
        for (int v = tid; v < P.V; v += blockDim.x) dst[v] = scratch ? scratch[v] : ldcg_dom(&P.dom[v]);
#ifdef PCP_SET
        {  // the label slot holds the bit sets behind the bounds
          uint32_t* dbits = reinterpret_cast<uint32_t*>(dst + P.V);
          const long long nw = (long long)P.V * P.bits_W;
          for (long long i = tid; i < nw; i += blockDim.x) dbits[i] = __ldcg(&P.bits[i]);
        }
#endif
        __syncthreads();
        if (tid == 0) {
          const int2 meta = make_int2(bin_n, (int)trail_cnt);
          B.label_meta[slot] = meta;
          B.branches[nb] = make_int4(slot, var, val, 1);      // x > val, explored second
          B.branch_meta[nb] = meta;
          B.branches[nb + 1] = make_int4(slot, var, val, 0);  // x <= val, explored first
          B.branch_meta[nb + 1] = meta;
          L->n_labels = slot + 1;
          L->cur_label = slot;
          L->n_branch = nb + 2;
          L->top = make_int4(slot, var, val, 0);
          L->top_meta = meta;
          L->top_valid = 1;
        }
      }
    }
    if (tid == 0) {
      L->nodes = n + 1;
      L->iterations += iters;
      L->last_status = status;
      if (stop) L->stopped = 1;
      else if (status == 1) L->solutions += 1;
      else if (status == -1) L->failures += 1;
    }
    __syncthreads();
  }
  // ---- decide whether the burst goes on; pop the next branch
  if (tid == 0) {
    int run = 0, status = 0;
    const int nb = L->n_branch;
    if (L->err) status = 2;
    else if (L->stopped) status = 2;                                          // StopNode -> EndOfSearch
    else if (done > 0 && L->last_status == 1 && !B.all_solutions) status = 1; // OneSolution returns
    else if (!L->root_pending && nb == 0) status = B.all_solutions ? 2 : -1;  // tree exhausted
    else if (done >= B.node_budget) status = 0;                               // slice used up
    else run = 1;
    if (run && !L->root_pending) {
      if (L->top_valid) { s_pop = L->top; s_pop_meta = L->top_meta; }
      else { s_pop = __ldcg(&B.branches[nb - 1]); s_pop_meta = __ldcg(&B.branch_meta[nb - 1]); }
      L->top_valid = 0;
      L->n_branch = nb - 1;
    }
    L->run = run;
    if (!run) { bc->status = status; bc->cmd = 1; }
  }
  __syncthreads();
  if (!L->run) return;
  const bool root = L->root_pending != 0;
  __syncthreads();  // (everybody has read the flag thread 0 clears below)
  if (root) {
    if (tid == 0) {
      L->root_pending = 0; L->cur_label = -1; bc->inl_slot = -1; bc->bin_n = L->bin_n; bc->cmd = 0;
      bc->n_labels = L->n_labels; bc->n_branch = L->n_branch; bc->nodes = L->nodes;  // the other CTAs' mirrors
    }
    __syncthreads();
    return;
  }
  const int4 br = s_pop;
  const int2 meta = s_pop_meta;
  const int Lb = br.x;
  const int cur_label = L->cur_label;
  __syncthreads();  // (everybody has read what thread 0 resets below)
  if (cur_label != Lb) {
    // Snapshot::restore: domains <- label copy, re-activate the trail suffix (store.rs:319-323)
    const int2* src = B.stack + (long long)Lb * B.stack_stride;
    for (int v = tid; v < P.V; v += blockDim.x) P.dom[v] = __ldcg(&src[v]);
#ifdef PCP_SET
    {
      const uint32_t* sbits = reinterpret_cast<const uint32_t*>(src + P.V);
      const long long nw = (long long)P.V * P.bits_W;
      for (long long i = tid; i < nw; i += blockDim.x) P.bits[i] = __ldcg(&sbits[i]);
    }
#endif
    const unsigned cnt = __ldcg(&ctl->trail_cnt);
    for (unsigned i = (unsigned)meta.y + tid; i < cnt; i += blockDim.x) {
      unsigned ref = __ldcg(&P.trail[i]);
      unsigned fam = ref >> 29, slot = ref & kSlotMask;
      uint32_t* act = fam == F_NARY ? P.nary_active_w : P.fam[fam].active;
      atomicOr(&act[slot >> 5], 1u << (slot & 31));
    }
    __syncthreads();
    if (tid == 0) ctl->trail_cnt = (unsigned)meta.y;
  }
  if (tid == 0) {
    const int slot = meta.x;  // propagators are truncated to the labelled length (store.rs:320)
    if (slot >= B.bin_cap) { L->err = 1; bc->status = 2; bc->cmd = 1; L->run = 0; }
    else {
      int4 d = br.w == 0 ? make_int4((int)((B_LESS << 28) | (unsigned)br.y), 0, -1, br.z + 1)    // x <= val
                         : make_int4((int)((B_LESS << 28) | kConstVar28), br.z, br.y, 0);         // val < x
      P.fam[F_BIN].tdesc[slot] = d;
      atomicOr(&P.fam[F_BIN].active[slot >> 5], 1u << (slot & 31));
      bc->inl_desc = d;
      bc->inl_slot = slot;
      bc->bin_n = slot + 1;
      bc->cmd = 0;
      L->bin_n = slot + 1;
      L->n_labels = Lb + 1;
      L->cur_label = -1;
      bc->n_labels = L->n_labels; bc->n_branch = L->n_branch; bc->nodes = L->nodes;  // the other CTAs' mirrors
    }
  }
  __syncthreads();
}

// A slice of one engine's device-resident search.  `P`, `B`: the launch parameters where the hot
// loops read them (the kernel's parameter space, or shared memory in a batched launch); `s_P`,
// `s_B`: their copies in shared memory for the out-of-line sections.
template <bool SMEM>
__device__ __forceinline__ void burst_launch_body(const Params& P, const BurstParams& B, const Params& s_P, const BurstParams& s_B,
                                                  char* smem) {
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
  __shared__ int s_flags[2];
  __shared__ int s_cmd, s_slot, s_bin_n;
  __shared__ InlineProp s_inl;
  __shared__ BurstLocal s_local;
  const Params* PS = &s_P;
  CtaState st;
  cta_init(P, st, smem, s_full, s_empty, s_flags, SMEM);
  if (st.dbm) for (int w = threadIdx.x; w < P.dirty_words; w += blockDim.x) st.dbm[w] = 0u;  // (ordered by the barrier below)
  Control* ctl = P.ctl;
  BurstCtl* bc = B.bc;

  if (threadIdx.x == 0) {
    s_flags[0] = s_flags[1] = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (!B.incremental) pre_issue(P, st);  // descriptors never depend on the node
  } else if (threadIdx.x == 32) {
    if (cta_rank() == 0) burst_load(bc, &s_local);
  }
  __syncthreads();
  st.gen = P.gen0;
  unsigned epoch = P.epoch0;

  // Every CTA mirrors the few scalars of the search state that decide whether the next step is
  // a plain descent to the left child (refreshed from `bc` at every device barrier below).
  __shared__ int s_mlabels, s_mbranch, s_mbin;
  __shared__ unsigned long long s_mnodes, s_sel[kWarps];
  __shared__ unsigned s_tc;
  unsigned long long done = 0;
  if (cta_rank() == 0) burst_host_step(s_P, s_B, &s_local, st.sdom, false, 0, 0, done);
  bool fast = false;  // uniform across the grid
  while (true) {
    if (!fast) {
      grid_barrier(P, st.gen, 0, false, s_flags, 0);  // the posted node is visible to every CTA
      if (threadIdx.x == 0) {
        s_cmd = *(volatile int*)&bc->cmd;
        s_slot = *(volatile int*)&bc->inl_slot;
        s_bin_n = *(volatile int*)&bc->bin_n;
        s_mlabels = *(volatile int*)&bc->n_labels;
        s_mbranch = *(volatile int*)&bc->n_branch;
        s_mnodes = *(volatile unsigned long long*)&bc->nodes;
        s_mbin = s_bin_n;
        if (s_slot >= 0) {
          s_inl.q[0] = __ldcg(&bc->inl_desc);
          s_inl.q[1] = s_inl.q[2] = make_int4(0, 0, 0, 0);
          s_inl.fam = F_BIN;
          s_inl.slot = s_slot;
        }
      }
      if (SMEM) for (int v = threadIdx.x; v < P.V; v += blockDim.x) st.sdom[v] = ldcg_dom(&P.dom[v]);
      __syncthreads();
      if (s_cmd != 0) break;
    }
    unsigned iters = 0;
    // Incremental search (PCP_FLAG_INCREMENTAL): a node = a label that was a fixpoint + one posted
    // constraint, so only that constraint and what it wakes are evaluated (its variable's row goes into
    // the worklist of iteration 0); the root schedules everything (store.rs:144-149).  No descriptors
    // are pre-issued then: the ring is the worklist's scratch memory.
    const bool sweep_first = !B.incremental || s_slot < 0;
    const unsigned dec = fixpoint_node<SMEM>(P, PS, st, epoch, s_bin_n, s_slot >= 0 ? 1 : 0, &s_inl, sweep_first, 0, !B.incremental, iters);
    epoch += iters + 1;
    ++done;
    // the next sweep's descriptors can stream in while CTA 0 does the host's work
    if (threadIdx.x == 0 && dec != D_ITER_CAP && !B.incremental) pre_issue(P, st);

    // Fast descent.  At a fixpoint every CTA's snapshot is the store (the last iteration narrowed
    // nothing), so every CTA can derive what CTA 0 is about to do: if the node is Unknown and
    // neither a limit nor a capacity stops the search, the next node is its left child --
    // FirstSmallestVar / MiddleVal on the snapshot, x <= val posted in the next tail slot
    // (binary_split.rs:46-57, explored first: one_solution.rs:46-51).  Then nobody waits for
    // CTA 0: the others apply the constraint to their snapshots and start the next sweep while
    // CTA 0 does the bookkeeping of burst_host_step (trace, label, branch records, descriptor).
    fast = false;
    unsigned long long best = ~0ull;
#if defined(PCP_SET) || defined(PCP_DEBUG_NO_FAST)
    constexpr bool kFastDescent = false;  // FirstSmallestVar compares cardinalities: one pass over the bit sets, by CTA 0 only
#else
    constexpr bool kFastDescent = true;
#endif
    if (kFastDescent && SMEM && dec == D_FIXPOINT) {
      if (threadIdx.x == 0) s_tc = __ldcg(&ctl->trail_at_decision);
      for (int v = threadIdx.x; v < P.V; v += blockDim.x) {
        const int2 d = st.sdom[v];
        const unsigned size = (unsigned)(d.y - d.x) + 1u;
        if (size > 1u) best = min(best, ((unsigned long long)size << 32) | (unsigned)v);
      }
      for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      if ((threadIdx.x & 31) == 0) s_sel[threadIdx.x >> 5] = best;
      __syncthreads();
      best = s_sel[0];
      for (int w = 1; w < kWarps; ++w) best = min(best, s_sel[w]);
      const bool unknown = (long long)s_tc != B.props_base + s_mbin;
      const bool stop = B.node_limit && s_mnodes + 1 >= B.node_limit;
      fast = unknown && best != ~0ull && !stop && done < B.node_budget && s_mlabels < B.max_labels &&
             s_mbranch + 2 <= B.max_branches && s_mbin < B.bin_cap;
    }
    if (!fast) dirty_sets_reset(P, st);  // (a device barrier follows before anybody marks a variable again)
    if (cta_rank() == 0) burst_host_step(s_P, s_B, &s_local, st.sdom, true, dec, iters, done);
    if (fast) {
      __syncthreads();  // everybody has read the mirrors (and CTA 0 is done with its snapshot)
      if (threadIdx.x == 0) {
        const int var = (int)(best & 0xffffffffu);
        const int2 d = st.sdom[var];
        const int val = (d.x + d.y) / 2;  // MiddleVal: truncating division
        s_inl.q[0] = make_int4((int)((B_LESS << 28) | (unsigned)var), 0, -1, val + 1);  // x <= val
        if (cta_rank() == 0) {  // what the bookkeeper posted must be what everybody derived
          const int4 posted = __ldcg(&bc->inl_desc);
          const int4 mine = s_inl.q[0];
          if (!s_local.run || *(volatile int*)&bc->inl_slot != s_mbin || posted.x != mine.x || posted.y != mine.y ||
              posted.z != mine.z || posted.w != mine.w)
            s_local.err = 3;
        }
        s_inl.q[1] = s_inl.q[2] = make_int4(0, 0, 0, 0);
        s_inl.fam = F_BIN;
        s_inl.slot = s_mbin;
        s_slot = s_mbin;
        s_bin_n = s_mbin + 1;
        s_mbin += 1;
        s_mlabels += 1;
        s_mbranch += 1;
        s_mnodes += 1;
      }
      __syncthreads();
    }
  }
  // drain the chunks that were pre-issued for a node that will not run in this launch
  if (threadIdx.x >> 5 > 0 && st.my_chunks > 0 && !B.incremental) {
    const int n = min(st.my_chunks, kStages);
    for (int i = 0; i < n; ++i) {
      const int q = st.pipe_pos + i, s = q % kStages;
      mbar_wait(&s_full[s], (q / kStages) & 1);
    }
  }
  if (cta_rank() == 0 && threadIdx.x == 0) {
    burst_store(bc, &s_local);
    ctl->epoch = epoch + 1;
    Result r;
    r.failed = 0;
    r.trail_cnt = *(volatile unsigned*)&ctl->trail_cnt;
    r.iterations = 0;
    r.epoch = epoch + 1;
    r.propagations = *(volatile unsigned long long*)&ctl->propagations;
    r.decision = 0;
    r.gen = st.gen;
    *P.result = r;
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(kThreads, 1) pcp_burst_kernel(const __grid_constant__ Params P,
                                                                const __grid_constant__ BurstParams B) {
  extern __shared__ __align__(128) char smem[];
  __shared__ Params s_P;  // copies of the launch parameters for the out-of-line cold sections
  __shared__ BurstParams s_B;
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(&P);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_P);
    for (int i = threadIdx.x; i < (int)(sizeof(Params) / 4); i += blockDim.x) dst[i] = src[i];
  }
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(&B);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_B);
    for (int i = threadIdx.x; i < (int)(sizeof(BurstParams) / 4); i += blockDim.x) dst[i] = src[i];
  }
  cta_place(blockIdx.x, gridDim.x);  // (its barrier also orders the copies above)
  burst_launch_body<SMEM>(P, B, s_P, s_B, smem);
}

// Slices of several engines' device-resident searches in ONE launch (pcp_search_step_many on
// searches whose engines run the same kernel variant with the same geometry): CTAs
// [k * group, (k + 1) * group) run engine k's search with batchP[k] / batchB[k].  The groups never
// meet: every search runs through its own node budget at its own pace.
__global__ void __launch_bounds__(kThreads, 1) pcp_burst_batch_kernel(const Params* __restrict__ batchP,
                                                                     const BurstParams* __restrict__ batchB, int group) {
  extern __shared__ __align__(128) char smem[];
  __shared__ Params s_P;
  __shared__ BurstParams s_B;
  const unsigned k = blockIdx.x / (unsigned)group;
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(batchP + k);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_P);
    for (int i = threadIdx.x; i < (int)(sizeof(Params) / 4); i += blockDim.x) dst[i] = __ldg(&src[i]);
  }
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(batchB + k);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_B);
    for (int i = threadIdx.x; i < (int)(sizeof(BurstParams) / 4); i += blockDim.x) dst[i] = __ldg(&src[i]);
  }
  cta_place(blockIdx.x - k * (unsigned)group, (unsigned)group);
  burst_launch_body<true>(s_P, s_B, s_P, s_B, smem);
}

