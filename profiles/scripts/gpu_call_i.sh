#!/bin/bash
# Last 1-GPU call of the session: the default bench line of the final code, and the parity suite once more.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary_i.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 400 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; stage bench $?
cut -c1-300 gpurun_out/bench_i.json >> $S
timeout 300 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_i.log 2>&1; stage pytest $?
tail -3 gpurun_out/pytest_gpu_i.log >> $S
cat $S
