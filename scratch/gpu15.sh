#!/bin/bash
# round-1 final measurements: benches (C2 default, C3, C4, C5), reference arm, launch list, ncu full (C2, C4, C5)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 200 python bench.py --workload c3 --steps 200 --warmup 10 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 400 python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
timeout 300 python bench.py --impl reference --steps 200 --warmup 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
for w in c2 c4 c3 c5; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    ed=d.get('e2e_device_search') or {}
    print('$w', 'ms/step', round(d['ms_per_step'],4), 'warm', round(d['warm']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'e2e_dev', round(ed.get('ms_per_step',0),4), 'frac', round(d['roofline']['frac'],4), 'iters', d['iterations_per_step'], 'props/step', d['propagations_per_step'], 'cpu', round(d['cpu_baseline']['value']/1e6,1))
except Exception as ex:
    print('$w', 'ERR', ex); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
tail -c 400 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v10.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
PCP_NO_BURST=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 20 -c 4 -f -o gpurun_out/prof_r1_c2_v10 python bench.py --steps 30 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu c2 rc=$?"
PCP_NO_BURST=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 6 -c 2 -f -o gpurun_out/prof_r1_c5_v10 python bench.py --workload c5 --steps 4 --warmup 3 > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu c5 rc=$?"
PCP_NO_BURST=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pcp_fixpoint -s 4 -c 2 -f -o gpurun_out/prof_r1_c4_v10 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/ncu_full_c4.log 2>&1; echo "ncu c4 rc=$?"
