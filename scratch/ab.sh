#!/bin/bash
# A/B of two builds on the same box: pcp_b200/libpcp_b200_base.so (A) vs pcp_b200/libpcp_b200.so (B)
cd $GRAFT_REPO_ROOT
for i in 1 2 3; do
  echo "A:"; PCP_B200_LIB=$PWD/pcp_b200/libpcp_b200_base.so timeout 100 python scratch/t9.py c2 2>&1 | head -2
  echo "B:"; timeout 100 python scratch/t9.py c2 2>&1 | head -2
done
