#!/bin/bash
# Round 2, GPU call G (1 GPU): the migrating-walk kernel with EIGHT shards on one device (RMAT-24, 4 rounds per batch): ncu of a
# mid-run super-step of the plain (not instrumented) kernel, and the timing.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2g_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mig_step_kernel -s 2000 -c 1 -o gpurun_out/r2_prof_mig8 -f \
    python profiles/run_migrate_local.py 24 4 8 > gpurun_out/r2g_mig_under_ncu.log 2>&1; stage ncu_mig8 $?
timeout 600 python profiles/run_migrate_local.py 24 4 8 > gpurun_out/r2_migrate_local_rmat24_w8.jsonl 2> gpurun_out/r2g.err; stage mig_local8 $?
cat gpurun_out/r2_migrate_local_rmat24_w8.jsonl >> $S
tail -3 gpurun_out/r2g.err >> $S
cat $S
