#!/bin/bash
# Round 2, GPU call T (4 GPUs): the N = 4 bench line with the default kernel variant (8 staged tuples per warp and destination beyond 4 shards).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2t_summary.txt
: > $S
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2t_bench_4gpu.json 2> gpurun_out/r2t_bench_4gpu.err
echo "rc=$?" >> $S
grep "bench " gpurun_out/r2t_bench_4gpu.err | tail -5 >> $S
tail -2 gpurun_out/r2t_bench_4gpu.err >> $S
cat $S
