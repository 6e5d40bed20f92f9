#!/bin/bash
cd $GRAFT_REPO_ROOT
for i in 1 2; do
  echo "A base (f94e4f9):"; PCP_B200_LIB=$PWD/pcp_b200/libpcp_b200_base.so timeout 200 python scratch/t9.py c2 2>&1 | head -2
  echo "B before outlining:"; PCP_B200_LIB=$PWD/pcp_b200/libpcp_b200_cur.so timeout 200 python scratch/t9.py c2 2>&1 | head -2
  echo "C outlined:"; timeout 200 python scratch/t9.py c2 2>&1 | head -2
done
timeout 900 python -m pytest tests -m gpu -x -q -k "not 5000" 2>&1 | tail -3
