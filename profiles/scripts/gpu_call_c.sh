#!/bin/bash
# Third gpurun call of the session: weighted alias-fold (parity + C3 throughput), wide-gather probe, default bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary_c.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_c.log 2>&1; stage pytest $?
tail -5 gpurun_out/pytest_gpu_c.log >> $S
timeout 120 profiles/probes/gather_probe > gpurun_out/gather_probe_wide.txt 2>&1; stage probe $?
cat gpurun_out/gather_probe_wide.txt >> $S
timeout 500 python profiles/run_configs.py > gpurun_out/configs_c.jsonl 2> gpurun_out/configs_c.err; stage configs $?
cat gpurun_out/configs_c.jsonl >> $S
timeout 400 python bench.py --scale 24 --weighted 1 --steps 5 --warmup 3 > gpurun_out/bench_c3_wfold.json 2> gpurun_out/bench_c3_wfold.err; stage bench_c3 $?
timeout 600 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; stage bench $?
cat $S
