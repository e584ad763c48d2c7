#!/bin/bash
# One gpurun call: parity tests, smoke, kernel A/B, bench line, launch list, one full ncu capture, text-I/O figures.
# Everything lands in gpurun_out/; each stage has its own timeout so that one hang cannot eat the box.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary.txt
: > $S
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv >> $S 2>&1
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; stage pytest $?
tail -5 gpurun_out/pytest_gpu.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; stage smoke $?
timeout 400 python profiles/run_ab.py 26 > gpurun_out/fold_ab.jsonl 2> gpurun_out/fold_ab.err; stage ab $?
cat gpurun_out/fold_ab.jsonl >> $S
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; stage bench $?
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_v5.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; stage ncu_list $?
timeout 700 ncu --set full --clock-control none --import-source on -k regex:walk_fold_conv -s 1 -c 1 -o gpurun_out/prof_v5 -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1; stage ncu_full $?
timeout 400 python profiles/run_io.py 22 > gpurun_out/text_io.jsonl 2> gpurun_out/text_io.err; stage io $?
cat gpurun_out/text_io.jsonl >> $S
cat $S
