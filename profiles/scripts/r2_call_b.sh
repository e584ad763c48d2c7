#!/bin/bash
# Round 2, GPU call B (1 GPU): parity of the migrating-walker kernel (all shards on one device) + the older sharded tests,
# then what the super-step machinery costs on one GPU next to the single-GPU kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2b_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 900 python -m pytest tests/test_gpu_migrate.py tests/test_gpu_sharded.py -q -x --timeout 300 > gpurun_out/r2b_pytest.log 2>&1; stage pytest $?
tail -15 gpurun_out/r2b_pytest.log >> $S
timeout 600 python profiles/run_migrate_local.py 22 4 > gpurun_out/r2_migrate_local.jsonl 2> gpurun_out/r2_migrate_local.err; stage migrate_local $?
cat gpurun_out/r2_migrate_local.jsonl >> $S
tail -5 gpurun_out/r2_migrate_local.err >> $S
cat $S
