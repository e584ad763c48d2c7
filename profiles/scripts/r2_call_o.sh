#!/bin/bash
# Round 2, GPU call O (1 GPU): the -m gpu suite with the parallel Vose build; build profiles of C3 (RMAT-24 weighted) and weighted C5.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2o_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2o_pytest.log 2>&1; stage pytest $?
tail -30 gpurun_out/r2o_pytest.log >> $S
timeout 600 python profiles/run_build_profile.py > gpurun_out/r2o_build_profiles.jsonl 2> gpurun_out/r2o_build_profiles.err; stage build_profiles $?
cat gpurun_out/r2o_build_profiles.jsonl >> $S
tail -3 gpurun_out/r2o_build_profiles.err >> $S
cat $S
