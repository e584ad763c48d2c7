#!/bin/bash
# Two-GPU call: multi-rank parity (tuple exchange + peer-gather with the v5 PEER kernel over real NVLink), then the bench line at N=2.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/summary_d.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_sharded_check.py > gpurun_out/dist_check_2gpu.log 2>&1; stage dist_check $?
tail -4 gpurun_out/dist_check_2gpu.log >> $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; stage bench2 $?
cut -c1-1500 gpurun_out/bench_2gpu.json >> $S
cat $S
