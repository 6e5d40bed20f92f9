#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
for w in c2 c4; do python - <<PY
import json
d=json.load(open('gpurun_out/bench_$w.json'))
ed=d.get('e2e_device_search') or {}
print('$w', 'ms/step', round(d['ms_per_step'],4), 'warm', round(d['warm']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'e2e_dev', round(ed.get('ms_per_step',0),4), 'frac', round(d['roofline']['frac'],4))
PY
done
