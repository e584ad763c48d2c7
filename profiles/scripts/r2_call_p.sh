#!/bin/bash
# Round 2, GPU call P (1 GPU): the -m gpu suite and the default bench with the one-sort build; then ncu --set full on two mid-run
# launches of mig_step_kernel<.., VCUT> (8 shards on one device, RMAT-24, 3 rounds, hub rows 0.5): what the sharded step kernel is bound by.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2p_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2p_pytest.log 2>&1; stage pytest $?
tail -8 gpurun_out/r2p_pytest.log >> $S
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; stage bench $?
python -c "
import json
d=json.load(open('gpurun_out/r2p_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['runs'], 'build', d['config']['build_s'], d['config']['build_ms_per_phase'], 'parity', d['parity_at_scale']['equal'], 'host_abi', d['e2e_host_abi']['value'])" >> $S 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mig_step_kernel -s 44 -c 2 -o gpurun_out/r2p_prof_mig_hub -f \
    python profiles/run_migrate_local.py 24 3 8 0.5 > gpurun_out/r2p_mig_under_ncu.log 2>&1; stage ncu_mig $?
tail -3 gpurun_out/r2p_mig_under_ncu.log >> $S
ls -la gpurun_out/*.ncu-rep >> $S 2>&1
cat $S
