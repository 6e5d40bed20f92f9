#!/usr/bin/env python
"""bench.py -- propagations/s and search-nodes/s of the propagation fixpoint (BASELINE.json).

A *step* is one round of search nodes, one per subtree context of the GPU (74 by default): each a
pass of the hot path (`Consistency::consistency`, reference src/libpcp/propagation/store.rs:247-257)
over the propagator store, at the node the reference's own search (OneSolution o Propagation o
Brancher(FirstSmallestVar, MiddleVal, BinarySplit), search/mod.rs:45-52) visits next in that
context's subtree; the K fixpoints of a round share one batched launch (pcp_consistency_batch).
Default workload: BASELINE configs[1], n-queens N=1000 (V=1000, P=1,498,500 XNeqY descriptors,
example/src/nqueens.rs:27-50).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --steps K --warmup W    # CPU arm (oracle port, all cores)

Numbers on the JSON line (our arm):
  value       propagations/s, device-timed (CUDA events on the lead engine's stream around the
              batched launch of each step, max over ranks), L2 flushed between steps
  warm        the same without the flush (the descriptors stay L2-resident between rounds)
  e2e         the same metric through the C ABI with host buffers: the C++ search driver calls
              pcp_restore / pcp_prop_alloc / pcp_consistency / pcp_domains_read / pcp_label per
              node and context (host threads pipelining their share of the contexts); wall clock,
              H2D of the posted descriptor and D2H of status + domains inside
  roofline    algorithmic bytes (32 B per binary propagation) / device time of the
              fixpoint launches (CUDA events, measured live in this run) vs the
              measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the oracle's flat variant, 1 thread, on the same first nodes of the same DFS
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_PROP = {"c2": 32, "c5": 32, "c4": 48, "c3": 32}  # SURVEY 8d "algorithmic bytes per propagation"
# what a sweep actually streams per propagation (the domains come from shared memory / L2):
# 16-byte binary descriptors, the compact 8-byte stream of the all-XNeqY stores on Interval domains (C2, C5),
# 24-byte ternary descriptors
STREAM_BYTES = {"c2": 8, "c5": 8, "c4": 24, "c3": 16}


def build_model(workload: str):
    from pcp_b200 import models
    if workload == "c2":
        return models.nqueens(1000), "n-queens N=1000 (V=1000, P=1498500 XNeqY), first DFS nodes of the reference search"
    if workload == "c5":
        return models.nqueens(5000), "n-queens N=5000 (V=5000, P=37492500 XNeqY), subtree-sharded DFS"
    if workload == "c3":
        return models.all_interval(500), "all-interval N=500 (V=999, 2 Distinct + 499 Disj2[XEqYPlusZ])"
    if workload == "c4":
        return models.random_arith_csp(), "random arithmetic CSP V=100000 P=1000000 XEqYPlusZ, single fixpoint"
    if workload.startswith("nq"):
        n = int(workload[2:])
        return models.nqueens(n), f"n-queens N={n}"
    raise SystemExit(f"unknown workload {workload}")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


class PyDfs:
    """The reference search node by node through the engine surface (used for the device-timed
    passes so that an L2 flush can be slotted between two rounds of nodes)."""

    def __init__(self, engine):
        from pcp_b200 import parallel
        self.e = engine
        self.par = parallel
        self.stack = []
        self.started = False

    def next_node(self):
        """Post the next node's branching constraint (host side) -- returns False when exhausted."""
        if not self.started:
            self.started = True
            return True
        if not self.stack:
            return False
        label, d = self.stack.pop()
        self.e.restore(label)
        self.par.post_decision(self.e, d)
        return True

    def after_fixpoint(self, status):
        if status != 0:
            return
        lo, hi = self.e.domains()
        var, val = self.par.select_branch(lo, hi)
        label = self.e.label()
        self.stack.append((label, (var, val, 1)))
        self.stack.append((label, (var, val, 0)))


def device_timed_pass(engines, steps, warmup, flush, torch, device, skip=0):
    """K engines side by side (K = len(engines)), each on its own subtree.  A step = one round:
    every engine's next DFS node is posted, then the K fixpoints are launched together
    (pcp_consistency_batch) and timed as one region with CUDA events on the engines' streams
    (fork after the optional L2 flush on the lead stream, join after the last kernel)."""
    import pcp_b200
    dfs = [PyDfs(e) for e in engines]
    flush_buf = torch.empty(int(os.environ.get("PCP_BENCH_FLUSH_MIB", "512")) << 20, dtype=torch.uint8, device=device) if flush else None
    # the flush runs on the lead engine's own stream, directly before the timed launches: the
    # events then bracket the kernels, not the host's launch latency on an idle stream
    ext = torch.cuda.ExternalStream(engines[0].cuda_stream(), device=device)
    torch.cuda.synchronize(device)
    props = iters = nodes = launches = 0
    ms = 0.0
    n = 0
    live = list(range(len(engines)))
    while n < skip + warmup + steps and live:
        live = [i for i in live if dfs[i].next_node()]
        if not live:
            break
        timed = n >= skip + warmup
        if flush and n >= skip:
            with torch.cuda.stream(ext):
                flush_buf.add_(1)
        sts, stats = pcp_b200.consistency_batch([engines[i] for i in live])
        if timed:
            props += sum(int(s.propagations) for s in stats)
            iters += sum(int(s.iterations) for s in stats)
            ms += float(stats[0].kernel_ms)   # the whole round: first launch to last completion
            nodes += len(live)
            launches += sum(int(s.launches) for s in stats)  # one launch per round when the batch shares one, else one per node
        for i, st in zip(live, sts):
            dfs[i].after_fixpoint(st)
        n += 1
    return {"nodes": nodes, "propagations": props, "iterations": iters, "ms": ms, "launches": launches,
            "rounds": max(n - skip - warmup, 0)}


def run_ours(args):
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per context's stream
    import torch
    import torch.distributed as dist
    from pcp_b200 import Engine, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    workload = args.workload or "c2"
    model, desc = build_model(workload)
    bpp = BYTES_PER_PROP.get(workload, 32)
    peak, peak_src = peaks()
    set_domains = args.domains == "set"
    if set_domains:
        desc += " -- IntervalSet<i32> domains (FDSpace, as example/src/nqueens.rs allocates them)"
    # contexts per GPU: engines side by side, each on its own subtree.  The streaming-bound
    # stores (C5: DRAM-bound sweep, C4: one fixpoint) gain nothing from it.
    multi = workload in ("c2", "c3") or workload.startswith("nq") or args.contexts > 1
    K = args.contexts if args.contexts > 0 else (74 if multi else (37 if workload == "c5" else 1))
    multi = multi or K > 1
    # the host-driven loop runs one host thread per context, the device-resident searches one blocked
    # thread: their best context counts differ from the device-timed rounds' (measured: DESIGN 6)
    K_e2e = max(1, min(K, args.e2e_contexts)) if multi else 1
    # host threads of the host-driven loop: the rank's share of the cores; each thread pipelines its
    # share of the contexts (pcp_search_step_many)
    os.environ.setdefault("PCP_SEARCH_THREADS", str(max(1, min(K_e2e, 6, (os.cpu_count() or 8) // max(world, 1) - 1))))
    # device-resident searches share one launch per slice: one CTA per search fills the GPU (stores whose
    # label stacks are small enough for 148 of them; else as many as the device-timed rounds use)
    # (a device search reserves its whole label stack up front: max_labels x V x 8 B on Interval engines,
    # x 136 B per variable for C2's bit sets on IntervalSet engines -- 2.2 GB per search)
    small = (workload in ("c2", "c3") or workload.startswith("nq")) and not set_domains
    K_dev = (args.dev_contexts or (148 if small and args.contexts == 0 else min(K, 20))) if multi else 1
    K_inc = (args.inc_contexts or (148 if small and args.contexts == 0 else min(K, 30))) if multi else K

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def reduce(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=op)
        return float(t.item())

    reduce_max = lambda x: reduce(x, dist.ReduceOp.MAX)  # noqa: E731
    reduce_sum = lambda x: reduce(x, dist.ReduceOp.SUM)  # noqa: E731

    single_fixpoint = workload == "c4"
    sampler = ClockSampler(local_rank) if rank == 0 else None
    V = model.num_vars
    sms = torch.cuda.get_device_properties(device).multi_processor_count

    def fresh_engine(mdl=None, host_search=False, timing=True, incremental=False, interval_set=None):
        e = Engine(device=local_rank, timing=timing, max_labels=1 << 14, host_search=host_search,
                   incremental=incremental, interval_set=set_domains if interval_set is None else interval_set)
        (mdl or model).load_into(e)
        return e

    def contexts(k, mdl=None, **kw):
        """k engines of this rank, each entered into its own frontier subtree (rank-major slices of
        the same breadth-first frontier on every rank)."""
        first = fresh_engine(mdl, **kw)
        parts = world * k
        paths = parallel.expand_frontier(first, parts=parts) if parts > 1 else [[]]
        if parts == 1:
            first.consistency()
        # the other contexts are forks of the first one at the root fixpoint: one upload and one reactor
        # build of the model, its static part shared on the device (pcp_engine_fork)
        engines = [first] + [first.fork() for _ in range(k - 1)]
        for i, e in enumerate(engines):
            if k > 1:
                e.set_grid_limit(max(1, sms // k))
            root = e.label()
            parallel.enter_subtree(e, root, paths[(rank * k + i) % len(paths)])
        return engines

    def close_all(engines):
        for e in engines:
            e.close()

    extra = {}
    if single_fixpoint:
        # C4: a step = one whole fixpoint from the initial domains (root label restored per step)
        e = fresh_engine()
        e.consistency()  # builds the CSR, warms up
        flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=device)
        res = {}
        for mode in ("flush", "warm"):
            e2 = fresh_engine()
            root = e2.label()
            ext = torch.cuda.ExternalStream(e2.cuda_stream(), device=device)
            torch.cuda.synchronize(device)
            props = iters = 0
            ms = 0.0
            barrier()
            for i in range(args.warmup + args.steps):
                e2.restore(root)
                if mode == "flush":
                    with torch.cuda.stream(ext):
                        flush_buf.add_(1)
                st, stats = e2.consistency()
                if i >= args.warmup:
                    props += stats.propagations
                    iters += stats.iterations
                    ms += stats.kernel_ms
            barrier()
            res[mode] = {"nodes": args.steps, "propagations": props, "iterations": iters, "ms": ms,
                         "launches": args.steps, "rounds": args.steps}
            e2.close()
        # e2e: restore + consistency + domains through the ABI, wall clock
        e3 = fresh_engine(timing=False)
        root = e3.label()
        e3.consistency()
        e3.restore(root)
        barrier()
        t0 = time.perf_counter()
        props_e2e = 0
        for i in range(args.steps):
            e3.restore(root)
            st, stats = e3.consistency()
            lo, hi = e3.domains()
            props_e2e += stats.propagations
        torch.cuda.synchronize(device)
        e2e_s = time.perf_counter() - t0
        e2e = {"propagations": props_e2e, "seconds": e2e_s, "nodes": args.steps}
        e2e_dev = None
        h2d, d2h = 0, 64 + 8 * V
        K = 1
    else:
        res = {}
        for mode in ("flush", "warm"):
            engines = contexts(K)
            barrier()
            res[mode] = device_timed_pass(engines, args.steps, args.warmup, mode == "flush", torch, device, skip=args.skip_nodes)
            barrier()
            close_all(engines)
        if K > 1:
            # one context with the whole GPU: the round-1 measurement, one node per step
            engines = contexts(1)
            barrier()
            res["single"] = device_timed_pass(engines, args.steps, args.warmup, True, torch, device, skip=args.skip_nodes)
            barrier()
            close_all(engines)
        # the same nodes with PCP_FLAG_INCREMENTAL: a node restored from a label that was a fixpoint
        # evaluates the posted constraint and what it wakes instead of scheduling every propagator
        # (store.rs:144-149) -- same domains and statuses (tests/), far fewer propagations
        engines = contexts(1, incremental=True)
        barrier()
        res["incremental"] = device_timed_pass(engines, args.steps, args.warmup, True, torch, device)
        barrier()
        close_all(engines)
        if K > 1:  # and the headline's own protocol (K contexts per round, L2 flushed) on incremental engines
            engines = contexts(K, incremental=True)
            barrier()
            res["incremental_batch"] = device_timed_pass(engines, args.steps, args.warmup, True, torch, device, skip=args.skip_nodes)
            barrier()
            close_all(engines)

        # e2e through the C++ driver over the C ABI with host buffers, twice: the host-driven node
        # loop (what a libpcp host does per node: restore, post one descriptor, fixpoint, status +
        # domains back, label) over the K contexts in lockstep is the `e2e` key; the device-resident
        # searches (same entry points, branching on the GPU) are reported beside it.  With more than
        # one rank the one-word collective (stop flag; sum all-reduce) runs inside the timed loop
        # after every round of `sync_every` nodes per context.
        def e2e_pass(host_search, k, incremental=False, steps=None):
            steps = steps or args.steps
            engines = contexts(k, host_search=host_search, timing=False, incremental=incremental)  # wall clock only: no event records, zero-copy results
            from pcp_b200 import search_step_many
            handles = [e.search_open(all_solutions=True) for e in engines]
            stop = parallel.StopFlag(device) if world > 1 else None
            search_step_many(handles, args.warmup)   # untimed: CSR, buffers, first nodes
            base = [(int(r.propagations), int(r.num_nodes)) for r in search_step_many(handles, 1)]
            barrier()
            sync_every = max(1, min(args.sync_every, steps)) if world > 1 else steps  # (one slice when there is nobody to talk to)
            done = exch = 0
            exch_s = 0.0
            t0 = time.perf_counter()
            last = None
            while done < steps:
                n = min(sync_every, steps - done)
                last = search_step_many(handles, n)
                done += n
                if stop is not None:
                    t1 = time.perf_counter()
                    stop.exchange(0, op="sum")
                    exch_s += time.perf_counter() - t1
                    exch += 1
            wall = time.perf_counter() - t0
            out = {"propagations": sum(int(r.propagations) - b[0] for r, b in zip(last, base)),
                   "nodes": sum(int(r.num_nodes) - b[1] for r, b in zip(last, base)), "seconds": wall,
                   "exchanges": exch, "exchange_seconds": exch_s, "contexts": k, "nodes_per_context": steps}
            barrier()
            for h in handles:
                h.close()
            close_all(engines)
            return out
        e2e = e2e_pass(True, K_e2e)
        e2e_dev = e2e_pass(False, K_dev, steps=max(args.steps, 200))  # (a 20-node window is two launches per context: too short to time)
        # PCP_FLAG_INCREMENTAL with the device-resident searches: nodes below the root evaluate their
        # posted constraint and the row of its variable instead of scheduling every propagator
        # (same statuses and domains: tests/); far fewer propagations per node, so nodes/s is its number
        e2e_inc = e2e_pass(False, K_inc, incremental=True, steps=max(args.steps, 400)) if multi else None
        h2d, d2h = 16 * K_e2e, (64 + 8 * V) * K_e2e  # per e2e step = one node on each of the K_e2e contexts: one posted descriptor in, header + domains out, each

        if world > 1 and workload == "c2" and not args.no_c5:
            extra["c5"] = run_c5(args, torch, dist, device, rank, world, local_rank, barrier, reduce_sum, reduce_max, peak)
            extra["branch_and_bound"] = run_bb(torch, device, rank, world, local_rank, barrier)

    clocks = sampler.stop() if sampler else None

    # ---- aggregate over ranks: units of all ranks / max time
    f = res["flush"]
    tot_props = reduce_sum(f["propagations"])
    tot_nodes = reduce_sum(f["nodes"])
    max_ms = reduce_max(f["ms"])
    w = res["warm"]
    w_props, w_nodes, w_ms = reduce_sum(w["propagations"]), reduce_sum(w["nodes"]), reduce_max(w["ms"])
    one = res.get("single")
    if one is not None:
        o_props, o_nodes, o_ms = reduce_sum(one["propagations"]), reduce_sum(one["nodes"]), reduce_max(one["ms"])
    inc = res.get("incremental")
    if inc is not None:
        i_props, i_nodes, i_ms = reduce_sum(inc["propagations"]), reduce_sum(inc["nodes"]), reduce_max(inc["ms"])
    incb = res.get("incremental_batch")
    if incb is not None:
        ib_props, ib_nodes, ib_ms = reduce_sum(incb["propagations"]), reduce_sum(incb["nodes"]), reduce_max(incb["ms"])
    e_props, e_nodes, e_s = reduce_sum(e2e["propagations"]), reduce_sum(e2e["nodes"]), reduce_max(e2e["seconds"])
    if e2e_dev is not None:
        d_props, d_nodes, d_s = (reduce_sum(e2e_dev["propagations"]), reduce_sum(e2e_dev["nodes"]),
                                 reduce_max(e2e_dev["seconds"]))
    e2e_inc = locals().get("e2e_inc")
    if e2e_inc is not None:
        n_props, n_nodes, n_s = (reduce_sum(e2e_inc["propagations"]), reduce_sum(e2e_inc["nodes"]),
                                 reduce_max(e2e_inc["seconds"]))
    launches = int(reduce_sum(f["launches"]))

    if rank == 0:
        # ---- CPU baseline: the oracle (flat variant), 1 thread, same first nodes of the same DFS
        cpu = cpu_baseline(workload, model, args, set_domains=set_domains)
        value = tot_props / (max_ms * 1e-3) if max_ms > 0 else 0.0
        achieved = (f["propagations"] * bpp) / (f["ms"] * 1e-3) / 1e9 if f["ms"] > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(workload)
            except Exception:
                traffic = None
        rounds = max(f["rounds"], 1)
        stream_b = 16 if (set_domains and workload in ("c2", "c5")) else STREAM_BYTES.get(workload, bpp)  # (IntervalSet: the 16-byte loop)
        line = {
            "metric": "propagations/s (and nodes/s) of the per-node propagation fixpoint",
            "value": value, "unit": "propagations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / rounds, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "l2": "flushed between steps (512 MiB read-modify-write on the lead engine's stream directly before each timed round)",
                       "step": (f"one round of {K} search nodes, one per subtree context: {K} Consistency::consistency fixpoints launched together "
                                "(pcp_consistency_batch), each on its own (vstore, cstore) pair" if K > 1 else
                                "one search node = one Consistency::consistency fixpoint"),
                       "contexts_per_gpu": K, "skip_nodes": args.skip_nodes,
                       "parallelism": (f"subtree sharding: {world} GPU(s) x {K} context(s)"),
                       "timing": "CUDA events per step (fork after the flush on the lead stream, join after the last context's kernel), summed; max over ranks"},
            "nodes_per_s": tot_nodes / (max_ms * 1e-3) if max_ms > 0 else 0.0,
            "us_per_node": 1e3 * max_ms * world / max(tot_nodes, 1),
            "propagations_per_node": f["propagations"] / max(f["nodes"], 1),
            "iterations_per_node": f["iterations"] / max(f["nodes"], 1),
            "warm": {"value": w_props / (w_ms * 1e-3) if w_ms > 0 else 0.0, "unit": "propagations/s",
                     "nodes_per_s": w_nodes / (w_ms * 1e-3) if w_ms > 0 else 0.0, "ms_per_step": w_ms / max(w["rounds"], 1),
                     "l2": "not flushed (descriptors L2-resident)"},
            "single_context": (None if one is None else {
                "value": o_props / (o_ms * 1e-3) if o_ms > 0 else 0.0, "unit": "propagations/s",
                "nodes_per_s": o_nodes / (o_ms * 1e-3) if o_ms > 0 else 0.0, "ms_per_step": o_ms / max(one["rounds"], 1),
                "roofline_frac": ((one["propagations"] * bpp) / (one["ms"] * 1e-3) / 1e9 / peak) if one["ms"] > 0 else None,
                "note": "one context with the whole GPU, one node per step, L2 flushed (the round-1 headline measurement)"}),
            "incremental": (None if inc is None else {
                "nodes_per_s": i_nodes / (i_ms * 1e-3) if i_ms > 0 else 0.0, "ms_per_step": i_ms / max(inc["rounds"], 1),
                "propagations_per_node": inc["propagations"] / max(inc["nodes"], 1),
                "note": "PCP_FLAG_INCREMENTAL, one context: no schedule-everything first sweep when the restored state "
                        "was a fixpoint; same domains and statuses, L2 flushed between steps",
                "batch": (None if incb is None else {
                    "contexts_per_gpu": K, "nodes_per_s": ib_nodes / (ib_ms * 1e-3) if ib_ms > 0 else 0.0,
                    "ms_per_step": ib_ms / max(incb["rounds"], 1), "us_per_node": 1e3 * ib_ms * world / max(ib_nodes, 1),
                    "propagations_per_node": ib_props / max(ib_nodes, 1),
                    "note": "the headline's protocol (one round of K nodes per step through pcp_consistency_batch, L2 flushed "
                            "before every round, CUDA events) on PCP_FLAG_INCREMENTAL engines: compare ms_per_step with the "
                            "line's own"})}),
            "e2e": {"value": e_props / e_s if e_s > 0 else 0.0, "unit": "propagations/s",
                    "nodes_per_s": e_nodes / e_s if e_s > 0 else 0.0, "us_per_node": 1e6 * e_s * world / max(e_nodes, 1),
                    "ms_per_step": 1e3 * e_s / max(args.steps, 1), "contexts_per_gpu": e2e.get("contexts", 1),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "exchanges": e2e.get("exchanges"), "exchange_seconds": e2e.get("exchange_seconds"),
                    "path": ("restore + pcp_consistency + pcp_domains_read per step through the C ABI (ctypes), L2 not flushed"
                             if single_fixpoint else
                             f"host-driven node loop over {K_e2e} context(s), one host thread each (pcp_search_step_many): C++ search "
                             "driver -> C ABI (pcp_restore / pcp_prop_alloc / pcp_consistency / pcp_domains_read / pcp_label per node "
                             "and context), wall clock, L2 not flushed" + ("; 1 x int32 NCCL all-reduce every round inside the timed loop" if world > 1 else ""))},
            "e2e_device_search": (None if e2e_dev is None else {
                "value": d_props / d_s if d_s > 0 else 0.0, "unit": "propagations/s",
                "nodes_per_s": d_nodes / d_s if d_s > 0 else 0.0, "us_per_node": 1e6 * d_s * world / max(d_nodes, 1),
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "contexts_per_gpu": e2e_dev.get("contexts"),
                "path": f"pcp_search_step_many over {K_dev} device-resident searches (branching, label/restore and the fixpoints on "
                        "the device; one launch per budget slice shared by all of them -- pcp_burst_batch_kernel, a group of CTAs "
                        "per search); counters copied back when a slice ends"}),
            "incremental_device_search": (None if e2e_inc is None else {
                "nodes_per_s": n_nodes / n_s if n_s > 0 else 0.0, "us_per_node": 1e6 * n_s * world / max(n_nodes, 1),
                "value": n_props / n_s if n_s > 0 else 0.0, "unit": "propagations/s",
                "propagations_per_node": n_props / max(n_nodes, 1), "contexts_per_gpu": e2e_inc.get("contexts"),
                "nodes_per_context": e2e_inc.get("nodes_per_context"),
                "note": "PCP_FLAG_INCREMENTAL + device-resident searches: below the root a node evaluates its posted constraint "
                        "and what it wakes (the reference schedules every propagator at every node, store.rs:144-149; statuses "
                        "and domains are the same).  Wall clock through pcp_search_step_many."}),
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "pcp_fixpoint_kernel (node prologue, TMA sweep, worklist iterations, label snapshot)"
                                   + (f"; the {K} contexts of a step share one launch (pcp_fixpoint_batch_kernel, {K} groups of CTAs)" if K > 1 else ""),
                         "dram_frac": ((traffic * f["launches"]) / (f["ms"] * 1e-3) / 1e9 / peak) if traffic and f["ms"] > 0 and peak else None,
                         "note": ("frac counts the algorithmic 32 B per propagation of every context; the contexts of a launch read "
                                  "ONE copy of the descriptors (HBM once, then L2) and keep their domains in shared memory, so frac "
                                  "can exceed 1: HBM moves `traffic` bytes per launch (dram_frac of the peak) and the launch is bound "
                                  "by the sweep's issue rate (profiles/r2_c2_batch_fixpoint_*)" if K > 1 else
                                  "frac: algorithmic bytes per propagation over the launch duration against the measured HBM peak"),
                         "algorithmic_bytes_per_propagation": bpp,
                         "streamed_bytes_per_propagation": stream_b,
                         "streamed_frac": ((f["propagations"] * stream_b) / (f["ms"] * 1e-3) / 1e9 / peak) if f["ms"] > 0 else None,
                         "warm_frac": ((w["propagations"] * bpp) / (w["ms"] * 1e-3) / 1e9 / peak) if w["ms"] > 0 else None},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_c5(args, torch, dist, device, rank, world, local_rank, barrier, reduce_sum, reduce_max, peak):
    """BASELINE configs[4]: n-queens N=5000 (V=5000, P=37,492,500; 600 MB of descriptors, 300 MB
    compact stream per GPU) as a parallel-subtree search over the ranks: one context per GPU (the
    sweep is DRAM-bound), device-timed nodes and the host-driven sharded search with the one-word
    NCCL exchange inside the timed loop."""
    from pcp_b200 import Engine, models, parallel
    m = models.nqueens(5000)
    steps, warm = min(args.steps, 10), 3

    def engine(**kw):
        e = Engine(device=local_rank, max_labels=1 << 10, **kw)
        m.load_into(e)
        paths = parallel.expand_frontier(e, parts=world)
        parallel.enter_subtree(e, e.label(), paths[rank % len(paths)])
        return e

    e = engine(timing=True)
    barrier()
    dev = device_timed_pass([e], steps, warm, True, torch, device)
    barrier()
    e.close()
    e = engine(timing=False, host_search=True)
    stop = parallel.StopFlag(device)
    h = e.search_open(all_solutions=True)
    h.step(warm)
    base = h.step(1)
    base = (int(base.propagations), int(base.num_nodes))
    barrier()
    t0 = time.perf_counter()
    exch_s, last = 0.0, None
    for _ in range(steps):
        last = h.step(1)
        t1 = time.perf_counter()
        stop.exchange(0, op="sum")
        exch_s += time.perf_counter() - t1
    wall = time.perf_counter() - t0
    barrier()
    h.close()
    e.close()
    props, nodes, ms = reduce_sum(dev["propagations"]), reduce_sum(dev["nodes"]), reduce_max(dev["ms"])
    e_props = reduce_sum(int(last.propagations) - base[0])
    e_nodes = reduce_sum(int(last.num_nodes) - base[1])
    e_s = reduce_max(wall)
    return {"workload": "n-queens N=5000 (V=5000, P=37492500 XNeqY), subtree-sharded DFS, one context per GPU",
            "value": props / (ms * 1e-3) if ms > 0 else 0.0, "unit": "propagations/s", "steps": steps, "warmup": warm,
            "nodes_per_s": nodes / (ms * 1e-3) if ms > 0 else 0.0, "ms_per_step": ms / max(steps, 1),
            "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s",
                         "streamed_bytes_per_propagation": 8,
                         "streamed_frac": (dev["propagations"] * 8) / (dev["ms"] * 1e-3) / 1e9 / peak if dev["ms"] > 0 else None,
                         "algorithmic_frac": (dev["propagations"] * 32) / (dev["ms"] * 1e-3) / 1e9 / peak if dev["ms"] > 0 else None,
                         "note": "the sweep streams the compact 8-byte descriptors from HBM and reads the domains from shared memory: "
                                 "the fraction against the bytes it must stream is the meaningful one"},
            "e2e": {"value": e_props / e_s if e_s > 0 else 0.0, "unit": "propagations/s", "nodes_per_s": e_nodes / e_s if e_s > 0 else 0.0,
                    "exchanges": steps, "exchange_seconds": exch_s,
                    "path": "host-driven node loop through the C ABI, 1 x int32 NCCL all-reduce after every node inside the timed loop"}}


def run_bb(torch, device, rank, world, local_rank, barrier):
    """BranchAndBound across ranks (search/branch_and_bound.rs:76-92 + SURVEY 8e): minimise the
    middle queen of n-queens N=32 on IntervalSet engines, subtrees sharded over the ranks, the
    incumbent all-reduced (min) after every round of 32 nodes and adopted by every rank."""
    from pcp_b200 import Engine, models, parallel
    n = 32
    e = Engine(device=local_rank, interval_set=True, host_search=True, max_labels=1 << 12)
    models.nqueens(n).load_into(e)
    flag = parallel.StopFlag(device)
    barrier()
    t0 = time.perf_counter()
    out = parallel.sharded_search(e, rank, world, node_budget=3000, sync_every=32, stop_flag=flag, bb_mode=1, bb_var=n // 2,
                                  parts_per_rank=2)
    wall = time.perf_counter() - t0
    barrier()
    e.close()
    return {"workload": f"n-queens N={n} on IntervalSet domains, minimise q[{n // 2}], node budget 3000 per rank",
            "incumbent": out["incumbent"], "nodes_rank0": out["nodes"], "solutions_rank0": out["solutions"],
            "exchanges": out["exchanges"], "exchange_seconds": out["exchange_seconds"], "wall_seconds": wall,
            "collective": "NCCL all-reduce, 1 x int32: sum (stop / idle word) + min (incumbent) per round"}


def cpu_baseline(workload, model, args, threads: int = 1, budget_s: float = 20.0, set_domains: bool = False,
                 faithful: bool = True):
    """Time the oracle's flat variant on a bounded sample of the same workload."""
    from oracle.oracle_api import FAITHFUL, FLAT, SET, OracleEngine
    dom = SET if set_domains else 0

    def one(nodes, out, i):
        e = OracleEngine(FLAT + dom)
        model.load_into(e)
        if workload == "c4":
            t0 = time.perf_counter()
            st, stats = e.consistency()
            out[i] = (int(stats.propagations), time.perf_counter() - t0, 1)
        else:
            r, _ = e.search(node_limit=nodes, all_solutions=True)
            out[i] = (int(r.propagations), float(r.seconds), int(r.num_nodes))
        e.close()

    # size the sample: ~30 M propagations/s/thread measured for binary propagators
    nodes = args.warmup + args.steps
    est = {"c2": 0.06, "c5": 1.5, "c3": 0.02}.get(workload, 0.05) * nodes
    if est > budget_s:
        nodes = max(2, int(nodes * budget_s / est))
    out = [None] * threads
    ts = [threading.Thread(target=one, args=(nodes, out, i)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    wall = time.perf_counter() - t0
    props = sum(o[0] for o in out)
    secs = max(o[1] for o in out)
    nn = sum(o[2] for o in out)
    faith = None
    if faithful and workload in ("c2", "c3") and threads == 1:
        # the reference's own algorithmic structure (BASELINE.md 3 "cpu-faithful"): reactor rebuilt at
        # every node with the duplicate-subscription scan (indexed_deps.rs:69-77, store.rs:125-149),
        # boxed view trees -- 3 nodes of the same DFS
        e = OracleEngine(FAITHFUL + dom)
        model.load_into(e)
        r, _ = e.search(node_limit=3, all_solutions=True)
        faith = {"value": r.propagations / r.seconds if r.seconds > 0 else 0.0, "nodes_per_s": r.num_nodes / r.seconds if r.seconds > 0 else 0.0,
                 "sample": "3 nodes, oracle faithful variant (per-node reactor rebuild incl. duplicate scan, boxed views), 1 thread"}
        e.close()
    return {"value": props / secs if secs > 0 else 0.0, "unit": "propagations/s", "cores": threads, "kind": "port",
            "faithful": faith,
            "nodes_per_s": nn / secs if secs > 0 else 0.0,
            "sample": (f"{out[0][2]} node(s) per thread of the same DFS, oracle flat variant (static CSR reactor, inline "
                       f"descriptors), {threads} thread(s), wall {wall:.1f}s incl. model load"),
            "host_cpus": os.cpu_count()}


def run_reference(args):
    """The reference's own CPU implementation of the path: libpcp is Rust and cannot be built
    in this image (no cargo/rustc), so this arm times the oracle port (flat variant) with every
    host thread it can use -- one independent DFS replica per thread (libpcp itself is
    single-threaded, so this is the generous reading)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload or "c2"
    model, desc = build_model(workload)
    threads = max(1, os.cpu_count() or 1)
    threads = min(threads, 64)
    cpu = cpu_baseline(workload, model, args, threads=threads, budget_s=40.0, set_domains=args.domains == "set", faithful=False)
    steps = args.steps
    line = {
        "impl": "reference",
        "metric": "propagations/s (and nodes/s) of the per-node propagation fixpoint",
        "value": cpu["value"], "unit": "propagations/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps, "warmup": args.warmup,
        "ms_per_step": (1e3 / cpu["nodes_per_s"]) if cpu["nodes_per_s"] else None,  # whole-job: all replicas
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": desc, "step": "one search node = one Consistency::consistency fixpoint"},
        "nodes_per_s": cpu["nodes_per_s"],
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "propagations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_JSON_FD = None


def emit(line: dict) -> None:
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line: whatever a library prints on file descriptor 1 (NCCL's
    # version banner, torchrun notices of child ranks) is sent to stderr instead
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="c2 (default) | c3 | c4 | c5 | nq<N>")
    ap.add_argument("--contexts", type=int, default=0, help="engines side by side per GPU in the device-timed rounds (0 = default: 74 for c2/c3 -- two CTAs each --, 37 for c5, 1 for c4)")
    ap.add_argument("--e2e-contexts", type=int, default=74, help="contexts of the host-driven e2e loop (shared by PCP_SEARCH_THREADS host threads)")
    ap.add_argument("--dev-contexts", type=int, default=0, help="contexts of the device-resident searches (0 = default)")
    ap.add_argument("--inc-contexts", type=int, default=0, help="contexts of the incremental device-resident searches (0 = default)")
    ap.add_argument("--domains", default="interval", choices=["interval", "set"], help="Interval<i32> (VStoreFD) or IntervalSet<i32> (FDSpace) domains")
    ap.add_argument("--skip-nodes", type=int, default=0, help="advance every context by this many DFS nodes before the timed window (deep nodes)")
    ap.add_argument("--sync-every", type=int, default=8, help="multi-GPU e2e: nodes per context between two exchanges of the stop word")
    ap.add_argument("--no-c5", action="store_true", help="multi-GPU: skip the nested C5 (N=5000) and branch-and-bound runs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
