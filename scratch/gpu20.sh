#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
wc -l gpurun_out/bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "ref n2 rc=$?"; wc -l gpurun_out/bench_n2_ref.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n2.json'))
ed=d.get('e2e_device_search') or {}
print('n_gpus', d['n_gpus'], 'value %.3g' % d['value'], 'nodes/s', round(d['nodes_per_s']), 'ms/step', round(d['ms_per_step'],4), 'e2e nodes/s', round(d['e2e']['nodes_per_s']), 'dev nodes/s', round(ed.get('nodes_per_s',0)), 'frac', round(d['roofline']['frac'],3))
r=json.load(open('gpurun_out/bench_n2_ref.json')); print('ref', r['impl'], '%.3g' % r['value'], r['ms_per_step'])
PY
