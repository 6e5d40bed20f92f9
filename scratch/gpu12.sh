#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python scratch/t12.py 2>&1 | tail -14
timeout 300 python scratch/t9.py c2 c4 2>&1 | tail -5
PCP_TRACE=1 PCP_NO_BURST=1 timeout 200 python scratch/t11.py 60 > gpurun_out/trace_it1.log 2>&1
