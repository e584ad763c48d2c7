#!/bin/bash
# Round 2, GPU call F (8 GPUs): the N = 8 and N = 4 bench lines of the sharded (migrating-walker) walk.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2f_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
for N in ${NS:-8 4}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps ${STEPS:-10} --warmup ${WARMUP:-3} > gpurun_out/r2f_bench_${N}gpu.json 2> gpurun_out/r2f_bench_${N}gpu.err; stage bench$N $?
  cat gpurun_out/r2f_bench_${N}gpu.json >> $S
  grep "bench " gpurun_out/r2f_bench_${N}gpu.err | tail -12 >> $S
  tail -3 gpurun_out/r2f_bench_${N}gpu.err >> $S
done
cat $S
