#!/bin/bash
cd $GRAFT_REPO_ROOT
for i in 1 2; do
for lib in lib_cur libpcp_b200; do
  echo "$lib:"; PCP_B200_LIB=$PWD/pcp_b200/$lib.so timeout 300 python scratch/t9.py c2 c4 c5 2>&1 | tail -5
done
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
