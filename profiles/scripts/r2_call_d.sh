#!/bin/bash
# Round 2, GPU call D (1 GPU): ncu on the migrating-walk kernel (2 shards on one device, RMAT-24), the whole -m gpu suite with the
# round-2 parity additions, the default N = 1 bench line (parity_at_scale, host-ABI e2e leg, gather ceiling from the probe).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2d_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mig_step_kernel -s 70 -c 2 -o gpurun_out/r2_prof_mig -f \
    python profiles/run_migrate_local.py 24 2 2 > gpurun_out/r2d_mig_under_ncu.log 2>&1; stage ncu_mig $?
timeout 600 python profiles/run_migrate_local.py 24 2 2 > gpurun_out/r2_migrate_local_rmat24.jsonl 2> gpurun_out/r2_migrate_local_rmat24.err; stage mig_local $?
cat gpurun_out/r2_migrate_local_rmat24.jsonl >> $S
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2d_pytest.log 2>&1; stage pytest $?
tail -12 gpurun_out/r2d_pytest.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; stage smoke $?
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; stage bench $?
cat gpurun_out/r2d_bench.json >> $S
tail -30 gpurun_out/r2d_bench.err >> $S
cat $S
