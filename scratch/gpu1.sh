#!/bin/bash
# session GPU call 1: sanity + timelines + benches + ncu
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
PCP_TRACE=1 PCP_NO_BURST=1 timeout 120 python scratch/t2.py > gpurun_out/trace_c2.log 2>&1; echo "trace rc=$?"
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 200 python bench.py --workload c3 --steps 200 --warmup 10 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 400 python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v3.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
PCP_NO_BURST=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 20 -c 6 -f -o gpurun_out/prof_r1_c2_v3 python bench.py --steps 30 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 3 -c 3 -f -o gpurun_out/prof_r1_c4_v3 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
cat gpurun_out/bench_c2.json | cut -c1-1500
tail -5 gpurun_out/bench_c5.err; cat gpurun_out/bench_c5.json | cut -c1-800
