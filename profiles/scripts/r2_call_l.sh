#!/bin/bash
# Round 2, GPU call L (1 GPU): migrate / lean tests, then the sharded kernel with every shard on ONE GPU at RMAT-24, 8 shards,
# hub replication 0 / 0.25 / 0.5 / 0.75 (what the migrations themselves cost, before NVLink), then the build profile of the lean build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2l_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 900 python -m pytest tests/test_gpu_migrate.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q --timeout 600 -x > gpurun_out/r2l_pytest.log 2>&1; stage pytest $?
tail -15 gpurun_out/r2l_pytest.log >> $S
timeout 900 python profiles/run_migrate_local.py 24 2 8 0,0.25,0.5,0.75 > gpurun_out/r2l_migrate_local_hubs.jsonl 2> gpurun_out/r2l_migrate_local.err; stage migrate_local $?
cat gpurun_out/r2l_migrate_local_hubs.jsonl >> $S
tail -5 gpurun_out/r2l_migrate_local.err >> $S
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/r2l_bench_short.json 2> gpurun_out/r2l_bench_short.err; stage bench_short $?
python -c "
import json
d=json.load(open('gpurun_out/r2l_bench_short.json'))
print('build', d['config']['build_s'], d['config']['build_ms_per_phase'], 'value', d['value'])" >> $S 2>&1
cat $S
