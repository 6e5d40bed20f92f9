#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python scratch/t12.py 2>&1 | tail -14
timeout 300 python scratch/t9.py c2 c4 2>&1 | tail -4
timeout 100 python bench.py --workload c3 --steps 200 --warmup 10 | python -c "import json,sys; d=json.load(sys.stdin); print('c3 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'dev', d['e2e_device_search']['ms_per_step'])"
