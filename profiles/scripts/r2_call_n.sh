#!/bin/bash
# Round 2, GPU call N (2 GPUs): two-GPU tests (NCCL check incl. the VCut shard map, one process / two GPUs through the ABI and the CLI
# with --partitioned true), then the N = 2 bench with replicated hub rows: 0.5 (full line), 0.75 and 0 (A/B on the same box, value only).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2n_summary.txt
: > $S
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
CUDA_VISIBLE_DEVICES=0,1 timeout 500 python -m pytest tests/test_gpu_migrate.py -q -x --timeout 400 -k "two_ranks or one_process" > gpurun_out/r2n_pytest.log 2>&1; stage pytest2 $?
tail -6 gpurun_out/r2n_pytest.log >> $S
run() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 10 --warmup 3 $2 > gpurun_out/r2n_bench_2gpu_$1.json 2> gpurun_out/r2n_bench_2gpu_$1.err; stage bench8_$1 $?
  grep "bench " gpurun_out/r2n_bench_2gpu_$1.err | tail -5 >> $S
  tail -2 gpurun_out/r2n_bench_2gpu_$1.err >> $S
}
run hub50 "--hub-fraction 0.5" 29561
run hub75 "--hub-fraction 0.75 --no-e2e --no-parity" 29562
run hub00 "--hub-fraction 0 --no-e2e --no-parity" 29563
cat $S
