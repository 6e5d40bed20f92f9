"""Subtree sharding of the search across GPUs / host workers.

One fixpoint is a single global computation and is never split; what shards is the search
tree: sibling subtrees are independent (a child is `restore(label)` + one branching
constraint, reference search/branching/branch.rs:51-55).  The root is expanded breadth-first
into at least `parts` open nodes; a node travels as its decision path
`[(var, val, alternative)]` and is rebuilt on the owning engine by replaying the path
(recomputation), so nothing but a few integers crosses ranks.  The only collective on the
path is a 1 x int32 all-reduce of the "solution found" / incumbent word (NCCL on GPUs,
gloo in the CPU tests), issued every `sync_every` nodes.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import models

Decision = Tuple[int, int, int]  # (var, val, alternative): alternative 0 = x <= val, 1 = x > val


def post_decision(engine, d: Decision) -> None:
    """BinarySplit's alternatives (search/branching/binary_split.rs:46-57)."""
    var, val, alt = d
    if alt == 0:
        engine.prop_alloc(models.X_LESS_Y, [[var, 0], [models.VAR_CONSTANT, val + 1]])  # x <= val
    else:
        engine.prop_alloc(models.X_LESS_Y, [[models.VAR_CONSTANT, val], [var, 0]])      # x > val


def select_branch(lo: np.ndarray, hi: np.ndarray) -> Tuple[int, int]:
    """FirstSmallestVar + MiddleVal (first_smallest_var.rs:30-39, middle_val.rs:25-27)."""
    size = hi.astype(np.int64) - lo.astype(np.int64) + 1
    cand = np.where(size > 1, size, np.iinfo(np.int64).max)
    var = int(np.argmin(cand))  # argmin returns the first minimum
    s = int(lo[var]) + int(hi[var])
    val = int(s / 2)            # truncating division, like Rust's `/` on i32
    return var, val


def expand_frontier(engine, parts: int, max_expansions: int = 4096) -> List[List[Decision]]:
    """Breadth-first expansion of the root into >= `parts` open nodes (decision paths).

    `engine` holds the loaded model at its root state.  Deterministic: every rank computes
    the same frontier and takes its own slice.  The engine is left at the root state."""
    st, _ = engine.consistency()
    root = engine.label()
    if st != 0:
        engine.restore(root)
        return [[]]
    frontier: List[List[Decision]] = [[]]
    closed: List[List[Decision]] = []
    n = 0
    while len(frontier) + len(closed) < parts and frontier and n < max_expansions:
        path = frontier.pop(0)
        engine.restore(root)
        for d in path:
            post_decision(engine, d)
        st, _ = engine.consistency()
        n += 1
        if st != 0:
            if st == 1:
                closed.append(path)  # a solution leaf still counts as a unit of work
            continue
        lo, hi = engine.domains()
        var, val = select_branch(lo, hi)
        frontier.append(path + [(var, val, 0)])
        frontier.append(path + [(var, val, 1)])
    engine.restore(root)
    return frontier + closed


def my_slice(paths: Sequence[List[Decision]], rank: int, world: int) -> List[List[Decision]]:
    """Static round-robin assignment of frontier nodes to ranks."""
    return [p for i, p in enumerate(paths) if i % world == rank]


def enter_subtree(engine, root_label: int, path: Sequence[Decision]) -> None:
    """Rebuild a frontier node on this engine: restore the root and replay its decisions."""
    engine.restore(root_label)
    for d in path:
        post_decision(engine, d)


class StopFlag:
    """The one-word collective of the path: max-reduce of a `found` flag (or min/max of an
    incumbent for BranchAndBound, search/branch_and_bound.rs:76-92) over all ranks."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self._dist = dist if dist.is_available() and dist.is_initialized() else None
        self._t = torch.zeros(1, dtype=torch.int32, device=device if device is not None else "cpu")

    def exchange(self, local_value: int, op: str = "max") -> int:
        """All-reduce one int32; returns the global value (identity without a process group)."""
        self._t.fill_(int(local_value))
        if self._dist is not None:
            ops = {"max": self._dist.ReduceOp.MAX, "min": self._dist.ReduceOp.MIN, "sum": self._dist.ReduceOp.SUM}
            self._dist.all_reduce(self._t, op=ops[op])
        return int(self._t.item())


def sharded_search(engine, rank: int, world: int, node_budget: int, sync_every: int = 64, stop_flag=None,
                   stop_on_solution: bool = False, warmup_nodes: int = 0, parts_per_rank: int = 4,
                   bb_mode: int = 0, bb_var: int = 0):
    """Run this rank's share of a sharded DFS.

    Every rank expands the same frontier, takes its round-robin slice and spends `node_budget`
    nodes on it, subtree after subtree, in rounds of `sync_every` nodes; after each round the
    one-word collective of the path runs (every rank executes the same number of rounds, so
    the collective always matches): the max of the `found` flag, and -- with `bb_mode` 1
    (minimise) / 2 (maximise) `bb_var` -- the min / max of the incumbent, which every rank then
    adopts (`set_incumbent`) so that its next nodes post `bb_var < best` like
    search/branch_and_bound.rs:76-92 does within one process.

    Returns a dict of counters.  `seconds` is the host wall clock of the search driver,
    `kernel_seconds` device time of the fixpoint launches, `exchange_seconds` the wall clock of
    the collectives, `wall_seconds` the whole round loop (driver + collectives) after the
    warm-up round(s): the end-to-end time of the sharded search."""
    import time
    paths = expand_frontier(engine, parts=max(world, 1) * parts_per_rank)
    mine = my_slice(paths, rank, world)
    root = engine.label()
    out = {"nodes": 0, "propagations": 0, "iterations": 0, "solutions": 0, "seconds": 0.0, "kernel_seconds": 0.0,
           "subtrees": 0, "frontier": len(paths), "stopped": False, "exchanges": 0, "exchange_seconds": 0.0,
           "wall_seconds": 0.0, "incumbent": None}
    keys = (("nodes", "num_nodes"), ("propagations", "propagations"), ("iterations", "iterations"),
            ("solutions", "num_solution"), ("seconds", "seconds"), ("kernel_seconds", "kernel_seconds"))
    rounds = (node_budget + sync_every - 1) // sync_every
    handle, prev, idx, found = None, None, 0, 0
    INF = 2**31 - 1
    best = None  # the global incumbent as of the last exchange
    t_wall = None
    for _ in range(rounds):
        if t_wall is None and out["nodes"] >= warmup_nodes:
            t_wall = time.perf_counter()
        round_used = 0
        while round_used < sync_every and out["nodes"] < node_budget:
            if handle is None:
                if idx >= len(mine):
                    break
                enter_subtree(engine, root, mine[idx])
                idx += 1
                handle = engine.search_open(all_solutions=not stop_on_solution, bb_mode=bb_mode, bb_var=bb_var,
                                            warmup_nodes=warmup_nodes if out["subtrees"] == 0 else 0)
                if bb_mode and best is not None:
                    handle.set_incumbent(best)
                prev = {k: 0 for k, _ in keys}
                out["subtrees"] += 1
            before = out["nodes"]
            res = handle.step(min(sync_every - round_used, node_budget - out["nodes"]))
            for k, attr in keys:
                cur = getattr(res, attr)
                out[k] += cur - prev[k]
                prev[k] = cur
            round_used += out["nodes"] - before
            if bb_mode and res.has_bb_value:
                v = int(res.bb_value)
                if best is None or (v < best if bb_mode == 1 else v > best):
                    best = v
            if res.status == 1 and stop_on_solution:
                found = 1
            if res.status != 0:
                handle.close()
                handle = None
            if found:
                break
        if stop_flag is not None:
            t0 = time.perf_counter()
            # one word carries both "some rank found a solution" and "every rank ran out of work"
            idle = 1 if (handle is None and idx >= len(mine)) or out["nodes"] >= node_budget else 0
            word = stop_flag.exchange(found + (idle << 16), op="sum")
            stop = (word & 0xffff) > 0
            all_idle = (word >> 16) >= max(world, 1)
            if bb_mode:
                if bb_mode == 1:
                    g = stop_flag.exchange(INF if best is None else best, op="min")
                    best = None if g == INF else g
                else:
                    g = stop_flag.exchange(-INF if best is None else best, op="max")
                    best = None if g == -INF else g
                if best is not None and handle is not None:
                    handle.set_incumbent(best)
                out["exchanges"] += 1
            out["exchanges"] += 1
            out["exchange_seconds"] += time.perf_counter() - t0
            if stop:
                out["stopped"] = True
                break
            if all_idle:
                break
        elif found:
            out["stopped"] = True
            break
    if t_wall is not None:
        out["wall_seconds"] = time.perf_counter() - t_wall
    if handle is not None:
        handle.close()
    out["incumbent"] = best
    return out


class SubtreeContexts:
    """K engines side by side on ONE GPU, each holding the model and searching its own subtree
    (SURVEY 8e applied inside a device).  One fixpoint of n-queens N=1000 leaves most of a B200
    idle between its phases (launch, prologue, device barriers, worklist iterations: 23 us per
    node of which the sweep is 4); K contexts on num_sms / K CTAs each overlap those phases of
    one node with the sweeps of the others, and forks of one engine share one launch per round
    (consistency_batch) or per slice of their searches (step).  Every context is an ordinary
    (vstore, cstore) pair behind the C ABI -- per-node results are exactly those of a lone
    engine on that subtree."""

    def __init__(self, make_engine, model, k: int, paths: Sequence[List[Decision]] = None, device_sms: int = 148,
                 fork: bool = True):
        """`make_engine()` creates the first engine; the others are `fork()`s of it at the root
        fixpoint (one upload and one reactor build of the model, static part shared on the
        device) unless `fork=False` (every context loads the model itself)."""
        first = make_engine()
        model.load_into(first)
        frontier = list(paths) if paths is not None else (expand_frontier(first, parts=k) if k > 1 else [[]])
        if paths is not None or k == 1:
            first.consistency()  # builds the reactor CSR, uploads the store
        self.engines = [first]
        for i in range(1, k):
            if fork and hasattr(first, "fork"):
                e = first.fork()
            else:
                e = make_engine()
                model.load_into(e)
                e.consistency()
            self.engines.append(e)
        self.paths = []
        for i, e in enumerate(self.engines):
            if k > 1:
                e.set_grid_limit(max(1, device_sms // k))
            root = e.label()
            path = frontier[i % len(frontier)]
            enter_subtree(e, root, path)
            self.paths.append(path)
        self.handles = None

    def open(self, **kw):
        self.handles = [e.search_open(**kw) for e in self.engines]
        return self.handles

    def step(self, max_nodes: int = 0):
        from .engine import search_step_many
        return search_step_many(self.handles, max_nodes)

    def close(self):
        if self.handles:
            for h in self.handles:
                h.close()
            self.handles = None
        for e in self.engines:
            e.close()
        self.engines = []
