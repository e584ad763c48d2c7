#!/bin/bash
# Round 2, GPU call R (1 GPU): variants of the sharded step kernel (resident blocks per SM x staged tuples per warp and destination)
# with 8 shards on one device, RMAT-24, 3 rounds, hub rows 0.5: does occupancy move the latency-bound kernel?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2r_summary.txt
: > $S
for v in ${VARIANTS:-4,16 4,8 5,8 6,8}; do
  echo "== variant $v" >> $S
  SRW_MIG_VARIANT=$v timeout 300 python profiles/run_migrate_local.py 24 3 ${WORLD:-8} 0.5 2> gpurun_out/r2r_err_$v.txt | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'migrate' in d['config'] and not d['instrumented']:
        print(d['steps_per_s'], d['ms'], d['super_steps'], d['checksum_equals_single_gpu'])" >> $S 2>&1
  tail -2 gpurun_out/r2r_err_$v.txt >> $S
done
cat $S
