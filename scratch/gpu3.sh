#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
PCP_NO_BURST=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 6 -c 3 -f -o gpurun_out/prof_r1_c5_v4 python bench.py --workload c5 --steps 4 --warmup 3 > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu c5 rc=$?"
PCP_NO_BURST=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 20 -c 4 -f -o gpurun_out/prof_r1_c2_v4 python bench.py --steps 30 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu c2 rc=$?"
