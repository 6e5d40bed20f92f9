#!/bin/bash
# re-entry call: tests + benches + launch list + ncu full of the current kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python scratch/t9.py 2>&1 | tail -12
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 200 python bench.py --workload c3 --steps 200 --warmup 10 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 400 python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
for w in c2 c4 c3 c5; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    print('$w', 'ms/step', round(d['ms_per_step'],4), 'warm', round(d['warm']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'iters', d['iterations_per_step'], 'props/step', d['propagations_per_step'])
except Exception as ex:
    print('$w', 'ERR', ex); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v5.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
PCP_NO_BURST=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 20 -c 4 -f -o gpurun_out/prof_r1_c2_v5 python bench.py --steps 30 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu c2 rc=$?"
PCP_NO_BURST=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 6 -c 2 -f -o gpurun_out/prof_r1_c5_v5 python bench.py --workload c5 --steps 4 --warmup 3 > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu c5 rc=$?"
tail -c 600 gpurun_out/bench_ref.json
