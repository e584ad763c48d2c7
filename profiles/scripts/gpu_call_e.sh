#!/bin/bash
# Fourth 1-GPU call: parity suite with alias-fold as the default sampler, smoke, the default bench line (now with the
# exact-sampler sample), one ncu capture of the weighted alias-fold kernel on BASELINE config C3's graph.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary_e.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_e.log 2>&1; stage pytest $?
tail -5 gpurun_out/pytest_gpu_e.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_e.log 2>&1; stage smoke $?
timeout 600 python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; stage bench $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_wfold_conv -s 1 -c 1 -o gpurun_out/prof_wfold_c3 -f \
    python bench.py --scale 24 --weighted 1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-exact > gpurun_out/bench_under_ncu_wfold.log 2>&1; stage ncu_wfold $?
cat $S
