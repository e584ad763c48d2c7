#!/bin/bash
# Round 2, GPU call H (8 GPUs): where does the N = 8 kernel time go?  The sharded walk as it is, without path stores
# (SRW_MIG_DEBUG=1) and with the path stores redirected to the local GPU (SRW_MIG_DEBUG=2); timing only (paths are wrong with 1/2).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2h_summary.txt
: > $S
N=${N:-8}
for D in 0 1 2; do
  SRW_MIG_DEBUG=$D timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$D bench.py --gpus $N --steps 5 --warmup 1 --no-parity --no-e2e > gpurun_out/r2h_bench_d$D.json 2> gpurun_out/r2h_bench_d$D.err
  echo "== debug $D rc=$?" >> $S
  python -c "
import json
d=json.load(open('gpurun_out/r2h_bench_d$D.json'))
print('value %.3e ms/step %.1f' % (d['value'], d['ms_per_step']), d['roofline']['super_step_profile'])
" >> $S 2>&1
done
cat $S
