#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PCP_NO_BURST=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 4 -c 2 -f -o gpurun_out/prof_r1_c4_v6 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/ncu_full_c4.log 2>&1; echo "ncu c4 rc=$?"
