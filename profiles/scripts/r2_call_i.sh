#!/bin/bash
# Round 2, GPU call I (8 GPUs): migrate parity on the device (all shards on GPU 0), 2-rank NCCL check, then the N = 8 bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2i_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
CUDA_VISIBLE_DEVICES=0,1 timeout 400 python -m pytest tests/test_gpu_migrate.py -q -x --timeout 300 > gpurun_out/r2i_pytest.log 2>&1; stage pytest $?
tail -4 gpurun_out/r2i_pytest.log >> $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2i_bench_8gpu.json 2> gpurun_out/r2i_bench_8gpu.err; stage bench8 $?
grep "bench " gpurun_out/r2i_bench_8gpu.err | tail -8 >> $S
tail -3 gpurun_out/r2i_bench_8gpu.err >> $S
cat $S
