#!/usr/bin/env python
"""bench.py -- propagations/s and search-nodes/s of the propagation fixpoint (BASELINE.json).

A *step* is one search node: one pass of the hot path (`Consistency::consistency`,
reference src/libpcp/propagation/store.rs:247-257) over the propagator store, at the node
the reference's own search (OneSolution o Propagation o Brancher(FirstSmallestVar, MiddleVal,
BinarySplit), search/mod.rs:45-52) visits next.  Default workload: BASELINE configs[1],
n-queens N=1000 (V=1000, P=1,498,500 XNeqY descriptors, example/src/nqueens.rs:27-50).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --steps K --warmup W    # CPU arm (oracle port, all cores)

Numbers on the JSON line (our arm):
  value       propagations/s, device-timed (CUDA events on the engine's stream around the node
              prologue + fixpoint kernel of each step, max over ranks), L2 flushed between steps
  warm        the same without the flush (descriptors, 24 MB, stay L2-resident between nodes)
  e2e         the same metric through the C ABI with host buffers: the C++ search driver calls
              pcp_restore / pcp_prop_alloc / pcp_consistency / pcp_domains_read / pcp_label per
              node; wall clock, H2D of the posted descriptor and D2H of status + domains inside
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
    passes so that an L2 flush can be slotted between two nodes)."""

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


def device_timed_pass(engine, steps, warmup, flush, torch, device):
    """K nodes timed with CUDA events on the engine stream (stats.kernel_ms covers the node
    prologue + the fixpoint kernel); optional L2 flush (write 512 MiB) between nodes."""
    dfs = PyDfs(engine)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=device) if flush else None
    # the flush runs on the engine's own stream, directly before the timed launch: the events
    # then bracket the kernel, not the host's launch latency on an idle stream
    ext = torch.cuda.ExternalStream(engine.cuda_stream(), device=device)
    torch.cuda.synchronize(device)
    props = iters = 0
    ms = 0.0
    launches = 0
    n = 0
    while n < warmup + steps:
        if not dfs.next_node():
            break
        if flush:
            with torch.cuda.stream(ext):
                flush_buf.add_(1)
        st, stats = engine.consistency()
        if n >= warmup:
            props += stats.propagations
            iters += stats.iterations
            ms += stats.kernel_ms
            launches += 1  # pcp_fixpoint_kernel (restore, posted constraint and label copy ride inside)
        dfs.after_fixpoint(st)
        n += 1
    return {"nodes": max(n - warmup, 0), "propagations": props, "iterations": iters, "ms": ms, "launches": launches}


def run_ours(args):
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
    workload = args.workload or ("c2" if world == 1 else "c2")
    model, desc = build_model(workload)
    bpp = BYTES_PER_PROP.get(workload, 32)
    peak, peak_src = peaks()

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def reduce_max(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    single_fixpoint = workload == "c4"
    sampler = ClockSampler(local_rank) if rank == 0 else None
    V = model.num_vars

    def fresh_engine(host_search=False, timing=True, incremental=False):
        e = Engine(device=local_rank, timing=timing, max_labels=1 << 16, host_search=host_search,
                   incremental=incremental)
        model.load_into(e)
        return e

    if single_fixpoint:
        # C4: a step = one whole fixpoint from the initial domains (root label restored per step)
        e = fresh_engine()
        e.consistency()  # builds the CSR, warms up
        flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=device)
        res = {}
        for mode in ("flush", "warm"):
            e2 = fresh_engine()
            lo0, hi0 = e2.domains()
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
                         "launches": args.steps}
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
    else:
        res = {}
        for mode in ("flush", "warm"):
            e = fresh_engine()
            if world > 1:
                # every rank works on its own subtree: replay the decision path of frontier node `rank`
                paths = parallel.expand_frontier(e, parts=world)
                root = e.label()
                parallel.enter_subtree(e, root, paths[rank % len(paths)])
            barrier()
            res[mode] = device_timed_pass(e, args.steps, args.warmup, mode == "flush", torch, device)
            barrier()
            e.close()
        # the same nodes with PCP_FLAG_INCREMENTAL: a node restored from a label that was a fixpoint
        # evaluates the posted constraint and what it wakes instead of scheduling every propagator
        # (store.rs:144-149) -- same domains and statuses (tests/), far fewer propagations
        e = fresh_engine(incremental=True)
        if world > 1:
            paths = parallel.expand_frontier(e, parts=world)
            root = e.label()
            parallel.enter_subtree(e, root, paths[rank % len(paths)])
        barrier()
        res["incremental"] = device_timed_pass(e, args.steps, args.warmup, True, torch, device)
        barrier()
        e.close()
        # e2e through the C++ driver over the C ABI, twice: the host-driven node loop (what a
        # libpcp host does: one launch, one posted descriptor in, status + domains out per node)
        # is the `e2e` key; the device-resident search (same C entry point, branching on the
        # GPU, results copied back when the search stops) is reported beside it
        def e2e_pass(host_search):
            e = fresh_engine(host_search=host_search, timing=False)  # wall clock only: no event records, zero-copy results
            stop = parallel.StopFlag(device) if world > 1 else None
            barrier()
            if world > 1:
                out = parallel.sharded_search(e, rank, world, node_budget=args.warmup + args.steps, sync_every=64,
                                              stop_flag=stop, warmup_nodes=args.warmup, parts_per_rank=1)
                res_ = {"propagations": out["propagations"], "seconds": out["seconds"], "nodes": out["nodes"] - args.warmup}
            else:
                r, _ = e.search(node_limit=args.warmup + args.steps, all_solutions=True, warmup_nodes=args.warmup)
                res_ = {"propagations": int(r.propagations), "seconds": float(r.seconds), "nodes": int(r.num_nodes) - args.warmup}
            barrier()
            e.close()
            return res_
        e2e = e2e_pass(True)
        e2e_dev = e2e_pass(False)
        h2d, d2h = 16, 64 + 8 * V  # one posted descriptor in, result header + domains out

    clocks = sampler.stop() if sampler else None

    # ---- aggregate over ranks: units of all ranks / max time
    f = res["flush"]
    tot_props = reduce_sum(f["propagations"])
    tot_nodes = reduce_sum(f["nodes"])
    max_ms = reduce_max(f["ms"])
    w = res["warm"]
    w_props, w_nodes, w_ms = reduce_sum(w["propagations"]), reduce_sum(w["nodes"]), reduce_max(w["ms"])
    inc = res.get("incremental")
    if inc is not None:
        i_props, i_nodes, i_ms = reduce_sum(inc["propagations"]), reduce_sum(inc["nodes"]), reduce_max(inc["ms"])
    e_props, e_nodes, e_s = reduce_sum(e2e["propagations"]), reduce_sum(e2e["nodes"]), reduce_max(e2e["seconds"])
    if e2e_dev is not None:
        d_props, d_nodes, d_s = (reduce_sum(e2e_dev["propagations"]), reduce_sum(e2e_dev["nodes"]),
                                 reduce_max(e2e_dev["seconds"]))
    launches = int(reduce_sum(f["launches"]))

    if rank == 0:
        # ---- CPU baseline: the oracle (flat variant), 1 thread, same first nodes of the same DFS
        cpu = cpu_baseline(workload, model, args)
        value = tot_props / (max_ms * 1e-3) if max_ms > 0 else 0.0
        achieved = (f["propagations"] * bpp) / (f["ms"] * 1e-3) / 1e9 if f["ms"] > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(workload)
            except Exception:
                traffic = None
        line = {
            "metric": "propagations/s (and nodes/s) of the per-node propagation fixpoint",
            "value": value, "unit": "propagations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "l2": "flushed between steps (512 MiB read-modify-write on the engine stream directly before each timed launch)",
                       "step": "one search node = one Consistency::consistency fixpoint",
                       "parallelism": f"subtree-sharding x{world}" if world > 1 else "single engine",
                       "timing": "CUDA events on the engine stream per step, summed; max over ranks"},
            "nodes_per_s": tot_nodes / (max_ms * 1e-3) if max_ms > 0 else 0.0,
            "propagations_per_step": f["propagations"] / max(f["nodes"], 1),
            "iterations_per_step": f["iterations"] / max(f["nodes"], 1),
            "warm": {"value": w_props / (w_ms * 1e-3) if w_ms > 0 else 0.0, "unit": "propagations/s",
                     "nodes_per_s": w_nodes / (w_ms * 1e-3) if w_ms > 0 else 0.0, "ms_per_step": w_ms / max(args.steps, 1),
                     "l2": "not flushed (descriptors L2-resident)"},
            "incremental": (None if inc is None else {
                "nodes_per_s": i_nodes / (i_ms * 1e-3) if i_ms > 0 else 0.0, "ms_per_step": i_ms / max(args.steps, 1),
                "propagations_per_step": inc["propagations"] / max(inc["nodes"], 1),
                "note": "PCP_FLAG_INCREMENTAL: no schedule-everything first sweep when the restored state was a "
                        "fixpoint; same domains and statuses, L2 flushed between steps"}),
            "e2e": {"value": e_props / e_s if e_s > 0 else 0.0, "unit": "propagations/s",
                    "nodes_per_s": e_nodes / e_s if e_s > 0 else 0.0, "ms_per_step": 1e3 * e_s * world / max(e_nodes, 1),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": ("restore + pcp_consistency + pcp_domains_read per step through the C ABI (ctypes), L2 not flushed"
                             if single_fixpoint else
                             "host-driven node loop: C++ search driver -> C ABI (pcp_restore / pcp_prop_alloc / "
                             "pcp_consistency / pcp_domains_read / pcp_label per node), wall clock, L2 not flushed")},
            "e2e_device_search": (None if e2e_dev is None else {
                "value": d_props / d_s if d_s > 0 else 0.0, "unit": "propagations/s",
                "nodes_per_s": d_nodes / d_s if d_s > 0 else 0.0, "ms_per_step": 1e3 * d_s * world / max(d_nodes, 1),
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "path": "pcp_search_run with the device-resident DFS (pcp_burst_kernel): branching, label/restore and "
                        "the fixpoints in one launch per budget slice; counters copied back when it stops"}),
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "pcp_fixpoint_kernel (node prologue, TMA sweep, worklist iterations, label snapshot)",
                         "algorithmic_bytes_per_propagation": bpp,
                         "warm_frac": ((w["propagations"] * bpp) / (w["ms"] * 1e-3) / 1e9 / peak) if w["ms"] > 0 else None},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(workload, model, args, threads: int = 1, budget_s: float = 20.0):
    """Time the oracle's flat variant on a bounded sample of the same workload."""
    from oracle.oracle_api import FLAT, OracleEngine

    def one(nodes, out, i):
        e = OracleEngine(FLAT)
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
    return {"value": props / secs if secs > 0 else 0.0, "unit": "propagations/s", "cores": threads, "kind": "port",
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
    cpu = cpu_baseline(workload, model, args, threads=threads, budget_s=40.0)
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
