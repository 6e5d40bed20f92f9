#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scratch/t9.py 2>&1 | tail -8
PCP_TRACE=1 PCP_NO_BURST=1 timeout 200 python scratch/t11.py 420 2>&1 | grep -E "NODE|iter " > gpurun_out/trace_deep.log
