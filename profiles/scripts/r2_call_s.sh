#!/bin/bash
# Round 2, GPU call S (8 GPUs): the N = 8 bench line with the default kernel variant (8 staged tuples per warp and destination beyond 4 shards).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2s_summary.txt
: > $S
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2s_bench_8gpu.json 2> gpurun_out/r2s_bench_8gpu.err
echo "rc=$?" >> $S
grep "bench " gpurun_out/r2s_bench_8gpu.err | tail -5 >> $S
tail -2 gpurun_out/r2s_bench_8gpu.err >> $S
cat $S
