#!/bin/bash
cd $GRAFT_REPO_ROOT
for lib in libpcp_b200_base lib_aab2a8e lib_dfa3aa2 lib_f0f9f5a lib_9421186 libpcp_b200; do
  echo "$lib:"; PCP_B200_LIB=$PWD/pcp_b200/$lib.so timeout 200 python scratch/t9.py c2 2>&1 | head -2
done
