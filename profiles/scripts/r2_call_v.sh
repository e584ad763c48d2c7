#!/bin/bash
# Round 2, GPU call V (1 GPU): final validation -- the -m gpu suite, smoke(), the default bench; ncu --set full on two mid-run launches of
# the DEFAULT sharded step kernel (mig_step_kernel<0,4,1,8,0>: 8 shards on one device, RMAT-24, hub rows 0.5, 8-tuple stages).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2v_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2v_pytest.log 2>&1; stage pytest $?
tail -8 gpurun_out/r2v_pytest.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; stage smoke $?
tail -2 gpurun_out/r2v_smoke.log >> $S
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; stage bench $?
python -c "
import json
d=json.load(open('gpurun_out/r2v_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['runs'], 'build', d['config']['build_s'], d['config']['build_ms_per_phase'], 'parity', d['parity_at_scale']['equal'])" >> $S 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:mig_step_kernel<\(bool\)0' -s 44 -c 2 -o gpurun_out/r2v_prof_mig_default -f \
    python profiles/run_migrate_local.py 24 3 8 0.5 > gpurun_out/r2v_mig_under_ncu.log 2>&1; stage ncu_mig $?
tail -2 gpurun_out/r2v_mig_under_ncu.log | cut -c1-300 >> $S
cat $S
