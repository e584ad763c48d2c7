#!/bin/bash
# Round 2, GPU call E (1 GPU): the trimmed migrating-walk kernel -- ncu (2 shards on one device, RMAT-24), timing for 2 and 8 shards
# on one device, the migrate parity tests.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2e_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 600 python -m pytest tests/test_gpu_migrate.py -q -x --timeout 300 > gpurun_out/r2e_pytest.log 2>&1; stage pytest $?
tail -5 gpurun_out/r2e_pytest.log >> $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mig_step_kernel -s 60 -c 1 -o gpurun_out/r2_prof_mig2 -f \
    python profiles/run_migrate_local.py 24 2 2 > gpurun_out/r2e_mig_under_ncu.log 2>&1; stage ncu_mig $?
timeout 600 python profiles/run_migrate_local.py 24 2 2 > gpurun_out/r2_migrate_local_rmat24_b.jsonl 2> gpurun_out/r2e_a.err; stage mig_local2 $?
cat gpurun_out/r2_migrate_local_rmat24_b.jsonl >> $S
timeout 600 python profiles/run_migrate_local.py 24 2 8 > gpurun_out/r2_migrate_local_rmat24_c.jsonl 2> gpurun_out/r2e_b.err; stage mig_local8 $?
cat gpurun_out/r2_migrate_local_rmat24_c.jsonl >> $S
cat $S
