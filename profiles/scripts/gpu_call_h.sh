#!/bin/bash
# Sixth 1-GPU call: parity suite (async rounds), smoke, the default bench line with overlapped finalisation and the repeated e2e region.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary_h.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_h.log 2>&1; stage pytest $?
tail -5 gpurun_out/pytest_gpu_h.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_h.log 2>&1; stage smoke $?
timeout 600 python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; stage bench $?
cut -c1-400 gpurun_out/bench_h.json >> $S
cat $S
