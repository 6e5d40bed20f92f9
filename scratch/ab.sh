#!/bin/bash
# A/B of two builds on the same box: pcp_b200/libpcp_b200_base.so (A) vs pcp_b200/libpcp_b200.so (B)
cd $GRAFT_REPO_ROOT
W="${1:-c2}"
for i in 1 2; do
  echo "A:"; PCP_B200_LIB=$PWD/pcp_b200/libpcp_b200_base.so timeout 200 python scratch/t9.py $W 2>&1 | tail -6
  echo "B:"; timeout 200 python scratch/t9.py $W 2>&1 | tail -6
done
