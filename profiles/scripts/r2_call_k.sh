#!/bin/bash
# Round 2, GPU call K (1 GPU): the -m gpu suite (lean build, VCut shard map on one device), then the default bench (lean build).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2k_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2k_pytest.log 2>&1; stage pytest $?
tail -25 gpurun_out/r2k_pytest.log >> $S
timeout 900 python bench.py > gpurun_out/r2k_bench_lean.json 2> gpurun_out/r2k_bench_lean.err; stage bench_lean $?
tail -c 3000 gpurun_out/r2k_bench_lean.json >> $S
tail -5 gpurun_out/r2k_bench_lean.err >> $S
cat $S
