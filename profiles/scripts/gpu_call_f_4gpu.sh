#!/bin/bash
# Four-GPU call: the bench line at N=4 (replicated CSR + sharded_c4 peer-gather over 4 symmetric-memory blocks + tuple exchange).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary_f.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
SRW_PEER_BUDGET_S=20 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; stage bench4 $?
cut -c1-300 gpurun_out/bench_4gpu.json >> $S
tail -5 gpurun_out/bench_4gpu.err >> $S
cat $S
