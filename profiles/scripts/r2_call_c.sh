#!/bin/bash
# Round 2, GPU call C (2 GPUs): the migrating walk over real NVLink peer memory -- parity under torchrun (NCCL barrier), then the
# N = 2 bench line (value = sharded walk, replicas + parity checksum beside it).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2c_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
nvidia-smi topo -m > gpurun_out/r2c_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_sharded_check.py > gpurun_out/r2c_dist_check.log 2>&1; stage dist_check $?
tail -4 gpurun_out/r2c_dist_check.log >> $S
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps ${STEPS:-5} --warmup ${WARMUP:-2} > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err; stage bench2 $?
cat gpurun_out/r2c_bench_2gpu.json >> $S
grep "bench " gpurun_out/r2c_bench_2gpu.err | tail -25 >> $S
tail -5 gpurun_out/r2c_bench_2gpu.err >> $S
cat $S
