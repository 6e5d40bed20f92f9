#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/bench_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err; echo "bench c5 n2 rc=$?"
tail -3 gpurun_out/bench_c5_n2.err | cut -c1-300
for f in bench_n2 bench_c5_n2; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$f.json'))
    ed=d.get('e2e_device_search') or {}
    print('$f', 'n_gpus', d['n_gpus'], 'value %.3g' % d['value'], 'nodes/s', round(d['nodes_per_s']), 'ms/step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'e2e nodes/s', round(d['e2e']['nodes_per_s']), 'dev nodes/s', round(ed.get('nodes_per_s',0)))
except Exception as ex:
    print('$f', 'ERR', ex)
PY
done
