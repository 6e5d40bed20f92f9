#!/bin/bash
cd $GRAFT_REPO_ROOT
for i in 1 2; do
for lib in lib_cur lib_exp lib_o2 lib_cicc2; do
  [ -f pcp_b200/$lib.so ] || continue
  echo "$lib:"; PCP_B200_LIB=$PWD/pcp_b200/$lib.so timeout 200 python scratch/t9.py c2 2>&1 | head -2
done
done
