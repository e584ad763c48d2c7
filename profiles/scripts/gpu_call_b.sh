#!/bin/bash
# Second gpurun call of the session: full parity suite with the convergent classic-alias kernel, secondary configs,
# the L2::64B load flavour under ncu (does it halve the DRAM traffic at equal speed?), the bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary_b.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_b.log 2>&1; stage pytest $?
tail -5 gpurun_out/pytest_gpu_b.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_b.log 2>&1; stage smoke $?
timeout 500 python profiles/run_configs.py > gpurun_out/configs_v5.jsonl 2> gpurun_out/configs_v5.err; stage configs $?
cat gpurun_out/configs_v5.jsonl >> $S
SRW_FOLD_VAR=1 timeout 700 ncu --set full --clock-control none --import-source on -k regex:walk_fold_conv -s 1 -c 1 -o gpurun_out/prof_v5_64B -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_64B.log 2>&1; stage ncu_64B $?
timeout 600 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; stage bench $?
timeout 300 python profiles/run_exact.py > gpurun_out/exact_v5.jsonl 2> gpurun_out/exact_v5.err; stage exact $?
cat $S
