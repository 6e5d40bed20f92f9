"""One round of K contexts through consistency_batch with PCP_TRACE: per-kernel exit times."""
import os, sys
os.environ["PCP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pcp_b200
from pcp_b200 import models, Engine, parallel
K = int(sys.argv[1]) if len(sys.argv) > 1 else 24
m = models.nqueens(1000)
first = Engine(device=0, timing=True, max_labels=1 << 14)
m.load_into(first)
paths = parallel.expand_frontier(first, parts=K)
engines = [first] + [first.fork() for _ in range(K - 1)]
for i, e in enumerate(engines):
    e.set_grid_limit(max(2, 148 // K))
    root = e.label()
    parallel.enter_subtree(e, root, paths[i % len(paths)])
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
ext = torch.cuda.ExternalStream(engines[0].cuda_stream(), device="cuda:0")
for r in range(4):
    with torch.cuda.stream(ext):
        flush.add_(1)
    sts, stats = pcp_b200.consistency_batch(engines)
    print("round", r, "ms", stats[0].kernel_ms, flush=True)
