#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python scratch/t12.py 2>&1 | grep -v "iters [3-9]" | tail -8
timeout 300 python scratch/t9.py c2 c5 2>&1 | tail -5
