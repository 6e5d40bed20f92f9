#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python scratch/t13.py 2>&1 | tail -4
PCP_LAUNCH=plain timeout 200 python scratch/t13.py 2>&1 | tail -4
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; tail -3 gpurun_out/bench_c2.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2.json'))
print('c2 ms/step', round(d['ms_per_step'],4), 'warm', round(d['warm']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'e2e_dev', round(d['e2e_device_search']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))
PY
