#!/bin/bash
# Fifth 1-GPU call: parity suite (exact sampler v4 = cert2 by default, directory input), exact-sampler A/B, default bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary_g.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_g.log 2>&1; stage pytest $?
tail -5 gpurun_out/pytest_gpu_g.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_g.log 2>&1; stage smoke $?
timeout 500 python profiles/run_exact.py > gpurun_out/exact_g.jsonl 2> gpurun_out/exact_g.err; stage exact $?
cut -c1-260 gpurun_out/exact_g.jsonl >> $S
timeout 600 python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; stage bench $?
cat $S
