#!/bin/bash
# N-way comparison of builds on one box: scratch/abn.sh lib_a lib_b ... (names under pcp_b200/, without .so)
cd $GRAFT_REPO_ROOT
for i in 1 2; do
for lib in "$@"; do
  [ -f pcp_b200/$lib.so ] || continue
  echo "$lib:"; PCP_B200_LIB=$PWD/pcp_b200/$lib.so timeout 200 python scratch/t9.py c2 2>&1 | head -3
done
done
